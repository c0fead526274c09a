"""CPU checks of the DGL / torchmetrics stand-ins against dense fp32 PyTorch references (tolerance 1e-5:
floating-point model math, outside the bit-exact data path)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _block(n_src=40, n_dst=12, n_e=150, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_src, (n_e,), generator=g, dtype=torch.int32)
    dst = torch.randint(0, n_dst, (n_e,), generator=g, dtype=torch.int32)
    return src, dst, n_src, n_dst


def test_sageconv_mean_matches_dense_reference():
    import legion_b200  # noqa: F401
    from legion_b200 import trainer
    from dgl.nn.pytorch import SAGEConv, GraphConv
    src, dst, n_src, n_dst = _block()
    blk = trainer.make_block(src, dst, n_src, n_dst)
    x = torch.randn(n_src, 16)
    A = torch.zeros(n_dst, n_src)
    A.index_put_((dst.long(), src.long()), torch.ones(len(src)), accumulate=True)       # multigraph adjacency
    for out_dim in (8, 32):
        conv = SAGEConv(16, out_dim, "mean")
        deg = A.sum(1).clamp(min=1).unsqueeze(1)
        want = x[:n_dst] @ conv.fc_self.weight.T + ((A @ x) / deg) @ conv.fc_neigh.weight.T + conv.bias
        assert torch.allclose(conv(blk, x), want, atol=1e-5, rtol=1e-5)
        gc = GraphConv(16, out_dim, allow_zero_in_degree=True)
        dout = A.sum(0).clamp(min=1)
        din = A.sum(1).clamp(min=1)
        want = ((A * dout.pow(-0.5)[None, :]) @ x @ gc.weight) * din.pow(-0.5)[:, None] + gc.bias
        assert torch.allclose(gc(blk, x), want, atol=1e-5, rtol=1e-5)


def test_sage_model_trains_and_accuracy_metric():
    import legion_b200  # noqa: F401
    from legion_b200 import trainer
    import torchmetrics
    torch.manual_seed(0)
    s1, d1, n9, n7 = _block(60, 30, 300, 1)
    s2, d2, _, n5 = _block(30, 10, 80, 2)
    x = torch.randn(n9, 12)
    y = torch.arange(n5) % 3
    model = trainer.SAGE(12, 16, 3, 2, dropout=0.0)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    losses = [float(trainer.train_step(model, opt, x, y, [(s1, d1), (s2, d2)], [(n9, n7), (30, n5)])) for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0]
    m = torchmetrics.Accuracy("multiclass", num_classes=3)
    model.eval()
    with torch.no_grad():
        acc = m(torch.softmax(model([trainer.make_block(s1, d1, n9, n7), trainer.make_block(s2, d2, 30, n5)], x), 1), y)
    assert 0.0 <= float(acc) <= 1.0 and float(m.compute()) == float(acc)


def test_spmm_aggregation_equals_index_add(monkeypatch):
    """LEGION_SHIM_SPMM=1 (CSR SpMM) must give the same sums and the same input gradient as index_add_ (fp32, 1e-5)."""
    import legion_b200  # noqa: F401
    from legion_b200 import trainer
    from dgl import heterograph
    src, dst, n_src, n_dst = _block(300, 90, 2500, 3)
    x = torch.randn(n_src, 20, requires_grad=True)
    g = torch.randn(n_dst, 20)
    res = []
    for flag in (False, True):
        monkeypatch.setattr(heterograph, "_USE_SPMM", flag)
        blk = trainer.make_block(src, dst, n_src, n_dst)
        out = blk.sum_messages(x)
        (gx,) = torch.autograd.grad(out, x, g)
        res.append((out.detach(), gx))
    assert torch.allclose(res[0][0], res[1][0], atol=1e-5, rtol=1e-5)
    assert torch.allclose(res[0][1], res[1][1], atol=1e-5, rtol=1e-5)
    empty = trainer.make_block(src[:0], dst[:0], n_src, n_dst)
    assert float(empty.sum_messages(x.detach()).abs().sum()) == 0.0
