#!/usr/bin/env python
"""Single-process, two-GPU run of the partitioned-cache gather for ncu (ncu must not wrap multi-rank commands): GPU 0
samples and gathers, half of the feature rows live in GPU 1's shard and are read with in-kernel P2P loads over NVLink.

    gpurun --gpus 2 -- 'python tools/nvlink_ncu.py > gpurun_out/nvl_plain.log 2>&1 && \
        ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,pcie__read_bytes.sum,dram__bytes_read.sum,gpu__time_duration.sum \
            -k regex:k_gather --clock-control none --csv --log-file gpurun_out/r02_nvlink_gather.csv python tools/nvlink_ncu.py'
Prints per-batch gather time (CUDA events) and the payload GB/s that crossed the link."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import legion_b200 as L
    n_nodes = int(os.environ.get("NVL_NODES", "20000000"))
    host_frac = float(os.environ.get("NVL_HOST_FRAC", "0"))      # > 0: that share of the rows is served from pinned host memory
    cfg = dict(L.synth.CONFIGS["C3"], n_nodes=n_nodes)
    N, D, B, fanout = cfg["n_nodes"], cfg["dim"], cfg["batch"], cfg["fanout"]
    L._lib.check(L.lib().lgn_enable_peer_access(2), "peer access")
    dev0 = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ds = L.synth.make_dataset(N, cfg["avg_deg"], D, n_class=cfg["n_class"], backend="torch", device=dev0,
                              dmin_fp=L.synth.calibrate_dmin(cfg["avg_deg"], N))
    r = L.Runner(N, D, B, fanout, device=0, part=0, rng_mode=L.RNG_PHILOX, rng_seed=42, enable_hotness=True, n_lanes=2)
    r.bind_topology(ds.indptr, ds.indices)
    train = ds.train_ids.contiguous()
    r.bind_seeds(L.MODE_TRAIN, train, ds.labels[train.long()].contiguous())
    s = torch.cuda.Stream(device=dev0)
    for step in range(16):
        r.batch_generate(L.MODE_TRAIN, B, step, stream=s.cuda_stream, pipe=step % 2)
        r.run_batch(with_features=False, is_presc=True, stream=s.cuda_stream)
    torch.cuda.synchronize()
    r.set_dedup_capacity(max(1, r.max_ids()))
    nh, _ = r.hotness()
    order = L.hot_order(nh)
    kg = 2
    n_cached = int(N * (1.0 - host_frac))
    cap = (n_cached + kg - 1) // kg
    slot_of = L.place(order, cap, kg)
    base = ds.features
    if host_frac > 0:
        base = L.MappedHostArray((N, D), np.float32)
        rows = max(1, (1 << 28) // (4 * D))
        for lo in range(0, N, rows):
            base.array[lo:lo + rows] = ds.features[lo:lo + rows].cpu().numpy()
    shard0 = L.fill_feature_shard(order, cap, kg, 0, ds.features, D)
    L._lib.check(L.lib().lgn_set_device(1), "set device")
    shard1 = L.DevArray((cap, D), np.float32)              # lives on GPU 1
    order1 = L.DevArray((N,), np.int32)
    L._lib.check(L.lib().lgn_copy_d2d(C.c_void_p(order1.ptr), C.c_void_p(order.ptr), C.c_int64(4 * N)), "copy order")
    feats1 = L.DevArray((N, D), np.float32) if N * D * 4 < 40e9 else None
    if feats1 is not None:
        L._lib.check(L.lib().lgn_copy_d2d(C.c_void_p(feats1.ptr), C.c_void_p(ds.features.data_ptr()), C.c_int64(4 * N * D)), "copy features")
        L.fill_feature_shard(order1, cap, kg, 1, feats1, D, out=shard1)
        L._lib.check(L.lib().lgn_device_synchronize(), "sync")
        feats1.free()
    L._lib.check(L.lib().lgn_set_device(0), "set device")
    r.bind_features(base)
    r.bind_feature_cache([shard0, shard1], slot_of, cap)
    r.set_epoch(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_b = int(os.environ.get("NVL_BATCHES", "6"))
    rows = 0
    for i in range(n_b + 2):
        if i == 2:
            r.tier_counts(reset=True, stream=s.cuda_stream)
            e0.record(s)
        r.batch_generate(L.MODE_TRAIN, B, i, stream=s.cuda_stream, pipe=i % 2)
        r.run_batch(with_features=True, stream=s.cuda_stream)
        r.wait_pipe(i % 2, stream=s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    tiers = r.tier_counts(stream=s.cuda_stream)
    ms = e0.elapsed_time(e1) / n_b
    print("rows per tier over %d batches [local, peer, host]: %s; %.3f ms/batch (one lane at a time); peer payload %.1f GB/s, host payload %.1f GB/s"
          % (n_b, tiers, ms, tiers[1] * D * 4 / n_b / (ms / 1e3) / 1e9, tiers[2] * D * 4 / n_b / (ms / 1e3) / 1e9))
    assert r.status(stream=s.cuda_stream) == 0


if __name__ == "__main__":
    main()
