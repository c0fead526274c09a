"""DGLBlock(gidx, (src_ntypes, dst_ntypes), etypes): message-flow block; dst nodes are a prefix of src nodes."""
import torch


class DGLBlock:
    def __init__(self, gidx, ntypes=(["_N"], ["_N"]), etypes=("_E",)):
        self._g = gidx
        self._in_deg = None

    def number_of_src_nodes(self):
        return self._g.num_src

    def number_of_dst_nodes(self):
        return self._g.num_dst

    num_src_nodes = number_of_src_nodes
    num_dst_nodes = number_of_dst_nodes

    def num_edges(self):
        return self._g.row.numel()

    def edges(self):
        return self._g.row, self._g.col

    def in_degrees(self):
        if self._in_deg is None:
            self._in_deg = torch.bincount(self._g.col, minlength=self._g.num_dst)
        return self._in_deg

    def out_degrees(self):
        return torch.bincount(self._g.row, minlength=self._g.num_src)

    def sum_messages(self, h_src):
        """out[v] = sum over edges (u -> v) of h_src[u]  (copy_u + sum)."""
        out = torch.zeros((self._g.num_dst, h_src.shape[1]), dtype=h_src.dtype, device=h_src.device)
        return out.index_add_(0, self._g.col, h_src.index_select(0, self._g.row))

    def to(self, device):
        self._g.row, self._g.col = self._g.row.to(device), self._g.col.to(device)
        return self
