"""DGLBlock(gidx, (src_ntypes, dst_ntypes), etypes): message-flow block; dst nodes are a prefix of src nodes."""
import os
import warnings

import torch

# LEGION_SHIM_SPMM=1: aggregate with one CSR SpMM (no E x D temporary, no atomics) instead of index_select +
# index_add_.  Same sums up to fp32 summation order; opt-in until it is timed on the GPU.
_USE_SPMM = os.environ.get("LEGION_SHIM_SPMM", "0") == "1"


class DGLBlock:
    def __init__(self, gidx, ntypes=(["_N"], ["_N"]), etypes=("_E",)):
        self._g = gidx
        self._in_deg = None
        self._csr = {}

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

    def _adjacency(self, dtype):
        """[num_dst x num_src] CSR with one unit entry per edge (duplicates kept: the blocks are multigraphs)."""
        if dtype not in self._csr:
            col, perm = torch.sort(self._g.col, stable=True)
            crow = torch.zeros(self._g.num_dst + 1, dtype=torch.int64, device=col.device)
            torch.cumsum(torch.bincount(col, minlength=self._g.num_dst), 0, out=crow[1:])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self._csr[dtype] = torch.sparse_csr_tensor(crow, self._g.row[perm], torch.ones(col.numel(), dtype=dtype, device=col.device),
                                                           size=(self._g.num_dst, self._g.num_src))
        return self._csr[dtype]

    def sum_messages(self, h_src):
        """out[v] = sum over edges (u -> v) of h_src[u]  (copy_u + sum)."""
        if _USE_SPMM and self._g.row.numel() > 0:
            return torch.sparse.mm(self._adjacency(h_src.dtype), h_src)
        out = torch.zeros((self._g.num_dst, h_src.shape[1]), dtype=h_src.dtype, device=h_src.device)
        return out.index_add_(0, self._g.col, h_src.index_select(0, self._g.row))

    def to(self, device):
        self._g.row, self._g.col = self._g.row.to(device), self._g.col.to(device)
        return self
