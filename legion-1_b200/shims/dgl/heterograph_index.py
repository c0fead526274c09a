"""create_unitgraph_from_coo(num_ntypes, num_src, num_dst, row, col, formats, row_sorted=..): a bipartite COO."""


class UnitGraphIndex:
    def __init__(self, num_src, num_dst, row, col):
        self.num_src, self.num_dst = int(num_src), int(num_dst)
        self.row, self.col = row.long(), col.long()      # row = message source (local index), col = destination


def create_unitgraph_from_coo(num_ntypes, num_src, num_dst, row, col, formats, row_sorted=False, col_sorted=False):
    assert num_ntypes == 2, "the Legion trainers only build bipartite blocks"
    return UnitGraphIndex(num_src, num_dst, row, col)
