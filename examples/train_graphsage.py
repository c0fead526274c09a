#!/usr/bin/env python
"""A trainer with the shape of the reference's legion_graphsage.py (one process per GPU, `ipc_service` for the
data, DGL-style blocks, SAGE model, Adam, accuracy metric) written against this repo only: it uses
legion-1_b200/ipc_service.py and, when the real packages are missing, the stand-ins in legion-1_b200/shims.
Start the server first (legion-1_b200/_build/legion or legion_server.py), then:
    python examples/train_graphsage.py --gpu 0 --features_num 100 --class_num 47 --epoch 2
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "legion-1_b200"))
sys.path.insert(0, ROOT)
import ipc_service  # noqa: E402  (the drop-in module, found on sys.path exactly like the reference's extension)
import legion_b200  # noqa: E402,F401
from legion_b200 import trainer  # noqa: E402

try:
    import torchmetrics
except ImportError:
    sys.path.append(os.path.join(ROOT, "legion-1_b200", "shims"))
    import torchmetrics


def run(args):
    dev = torch.device("cuda", args.gpu)
    torch.cuda.set_device(dev)
    ipc_service.initialize()
    train_steps, valid_steps, test_steps = ipc_service.get_steps()
    model = trainer.SAGE(args.features_num, args.hidden_dim, args.class_num, args.hops_num, args.drop_rate).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=args.learning_rate)

    def batch():
        ids, feats, labels, b1s, b1d, b2s, b2d = ipc_service.get_next(args.features_num)
        n1s, n1d, n2s, n2d = ipc_service.get_block_size()
        return feats, labels, [(b1s, b1d), (b2s, b2d)], [(n1s, n1d), (n2s, n2d)]

    for epoch in range(args.epoch):
        model.train()
        t0 = time.time()
        for _ in range(train_steps):
            feats, labels, coo, sizes = batch()
            loss = trainer.train_step(model, opt, feats, labels, coo, sizes)
            torch.cuda.synchronize()
            ipc_service.synchronize()
        epoch_s = time.time() - t0
        model.eval()
        metric = torchmetrics.Accuracy("multiclass", num_classes=args.class_num).to(dev)
        with torch.no_grad():
            for _ in range(valid_steps):
                feats, labels, coo, sizes = batch()
                blocks = [trainer.make_block(s, d, ns, nd) for (s, d), (ns, nd) in zip(coo, sizes)]
                metric(torch.softmax(model(blocks, feats), 1), labels.long())
                torch.cuda.synchronize()
                ipc_service.synchronize()
        print("Epoch:{}, Cost:{} s, Val Acc: {}, Loss: {}".format(epoch, epoch_s, float(metric.compute()), float(loss)), flush=True)
    model.eval()
    metric = torchmetrics.Accuracy("multiclass", num_classes=args.class_num).to(dev)
    with torch.no_grad():
        for _ in range(test_steps):
            feats, labels, coo, sizes = batch()
            blocks = [trainer.make_block(s, d, ns, nd) for (s, d), (ns, nd) in zip(coo, sizes)]
            metric(torch.softmax(model(blocks, feats), 1), labels.long())
            torch.cuda.synchronize()
            ipc_service.synchronize()
    print("Accuracy on test data: {}".format(float(metric.compute())), flush=True)
    ipc_service.finalize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser("Train GNN.")
    ap.add_argument("--class_num", type=int, default=47)
    ap.add_argument("--features_num", type=int, default=100)
    ap.add_argument("--hidden_dim", type=int, default=256)
    ap.add_argument("--hops_num", type=int, default=2)
    ap.add_argument("--drop_rate", type=float, default=0.5)
    ap.add_argument("--learning_rate", type=float, default=0.003)
    ap.add_argument("--epoch", type=int, default=2)
    ap.add_argument("--gpu", type=int, default=0)
    run(ap.parse_args())
