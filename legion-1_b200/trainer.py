"""A GraphSAGE consumer of the pipeline with the structure of the reference trainer
(pytorch_extension/legion_graphsage.py:36-89: SAGE(n_layers) of SAGEConv('mean'), ReLU + dropout between
layers, cross-entropy, Adam) used by bench.py to report the GraphSAGE epoch time and by the tests to prove
that batches delivered through `ipc_service` train.  Uses real DGL when installed, the stand-ins in shims/
otherwise.  The model math itself is ordinary PyTorch and is not part of the accelerated path."""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

try:
    import dgl  # noqa: F401
except ImportError:
    sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims"))
    import dgl  # noqa: F401
from dgl.heterograph import DGLBlock
from dgl.nn.pytorch import SAGEConv


class SAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, dropout=0.5):
        super().__init__()
        dims = [in_feats] + [n_hidden] * (n_layers - 1) + [n_classes]
        self.layers = nn.ModuleList([SAGEConv(dims[i], dims[i + 1], "mean") for i in range(n_layers)])
        self.dropout = nn.Dropout(dropout)

    def forward(self, blocks, x):
        h = x
        for i, (layer, block) in enumerate(zip(self.layers, blocks)):
            h = layer(block, h)
            if i != len(self.layers) - 1:
                h = self.dropout(F.relu(h))
        return h


def make_block(src, dst, n_src, n_dst):
    gidx = dgl.heterograph_index.create_unitgraph_from_coo(2, n_src, n_dst, src, dst, "coo", row_sorted=True)
    return DGLBlock(gidx, (["_N"], ["_N"]), ["_E"])


def train_step(model, opt, features, labels, blocks_coo, block_sizes):
    """blocks_coo / block_sizes: outermost layer first, as get_next / get_block_size return them."""
    blocks = [make_block(s, d, ns, nd) for (s, d), (ns, nd) in zip(blocks_coo, block_sizes)]
    logits = model(blocks, features)
    loss = F.cross_entropy(logits, labels.long())
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss
