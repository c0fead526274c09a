#!/usr/bin/env python
"""End-to-end throughput of the drop-in deployment at full size: the `legion` server binary (one process, all GPUs,
reference command line) publishing mini-batches over the reference wire format to one consumer process per GPU that
uses the drop-in `ipc_service` module exactly like the reference trainers (get_next / synchronize), minus the model.

    gpurun --gpus 8 --timeout 900 -- 'python tools/server_e2e.py --gpus 8 --config C3 --nodes 40000000 --agg-mode 3'

The dataset is generated with the synthetic generator of the bench, written in the reference's on-disk format to a
tmpfs directory, loaded by the server's own loader (pinned host memory + UVA), presampled, cached and served.
Prints one JSON line: batches/s, sampled edges/s and extracted-feature GB/s over all GPUs, timed on the consumer side
after a warm-up, plus the server's own phase log."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LEGION = os.path.join(ROOT, "legion-1_b200", "_build", "legion")


def client(dev, dim, warmup):
    import torch
    from legion_b200 import ipc_service
    torch.cuda.set_device(dev)
    ipc_service.initialize()
    tr, va, te = ipc_service.get_steps()
    epochs = int(os.environ["E2E_EPOCHS"])
    hops = int(os.environ.get("E2E_HOPS", "2"))
    total = (tr + va) * epochs + te
    edges = rows = 0
    t0 = None
    for g in range(total):
        if g == warmup:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            edges = rows = 0
        if hops == 2:
            ids, feats, labels, b1s, b1d, b2s, b2d = ipc_service.get_next(dim)
        else:                                            # additive k-hop consumer API (lp_sage shape)
            ids, feats, labels, blocks = ipc_service.get_next_k(dim, hops)
            b1s = blocks[0][0]
        edges += int(b1s.numel())
        rows += int(ids.numel())
        ipc_service.synchronize()
    dt = time.perf_counter() - (t0 if t0 is not None else time.perf_counter())
    ipc_service.finalize()
    print(json.dumps({"dev": dev, "batches": max(0, total - warmup), "seconds": dt, "edges": edges, "rows": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--nodes", type=int, default=0)
    ap.add_argument("--agg-mode", type=int, default=0, help="0/1/2/3 = 1/2/4/8 GPUs per NVLink clique (reference argv[2])")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--cache-gb", type=float, default=38.0)
    ap.add_argument("--dir", default="/dev/shm/lgn_e2e")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--trainer", default="", help="legion_graphsage | legion_gcn: consume with the reference's UNCHANGED trainer "
                    "(oracle/_ref/trainers/*.bin + its own ipc_service extension, oracle/_ref/ext) and report the epoch times it prints")
    ap.add_argument("--rng", default="philox")
    ap.add_argument("--client", type=int, default=-1)
    ap.add_argument("--dim", type=int, default=0)
    a = ap.parse_args()
    if a.client >= 0:
        return client(a.client, a.dim, a.warmup)

    import numpy as np
    import torch
    import legion_b200 as L
    from legion_b200 import dataset_io
    cfg = dict(L.synth.CONFIGS[a.config])
    if a.nodes:
        cfg["n_nodes"] = a.nodes
    N, D, B = cfg["n_nodes"], cfg["dim"], cfg["batch"]
    shutil.rmtree(a.dir, ignore_errors=True)
    data = os.path.join(a.dir, "data")
    os.makedirs(data)
    t0 = time.perf_counter()
    dev = torch.device("cuda", 0)
    ds = L.synth.make_dataset(N, cfg["avg_deg"], D, n_class=cfg["n_class"], backend="torch", device=dev,
                              dmin_fp=L.synth.calibrate_dmin(cfg["avg_deg"], N))
    host = L.synth.Dataset(n_nodes=N, n_edges=ds.n_edges, dim=D, n_class=cfg["n_class"],
                           **{k: getattr(ds, k).cpu().numpy() for k in ("indptr", "indices", "features", "labels", "train_ids", "valid_ids", "test_ids")},
                           dmin_fp=ds.dmin_fp, seed=ds.seed, backend="numpy")
    del ds
    torch.cuda.empty_cache()
    dataset_io.write_dataset(data, host)
    dataset_io.write_meta_config(a.dir, data, host, B, int(a.cache_gb * 1e9), a.epochs)
    n_edges = host.n_edges
    del host
    t_gen = time.perf_counter() - t0

    env = dict(os.environ, LEGION_RNG=a.rng, LEGION_FANOUT=",".join(map(str, cfg["fanout"])), E2E_EPOCHS=str(a.epochs),
               E2E_HOPS=str(len(cfg["fanout"])))
    log = open(os.path.join(a.dir, "server.log"), "w")
    t0 = time.perf_counter()
    srv = subprocess.Popen([LEGION, str(a.gpus), str(a.agg_mode)], cwd=a.dir, env=env, stdout=log, stderr=subprocess.STDOUT)
    try:
        while True:
            if srv.poll() is not None:
                raise RuntimeError("server exited early:\n" + open(os.path.join(a.dir, "server.log")).read()[-3000:])
            if "System is ready for serving" in open(os.path.join(a.dir, "server.log")).read():
                break
            time.sleep(0.5)
        t_ready = time.perf_counter() - t0
        if a.trainer:      # the reference's own trainer, unchanged, one process per GPU through its mp.spawn
            ref = os.path.join(ROOT, "oracle", "_ref")
            tenv = dict(env, PYTHONPATH=os.pathsep.join([os.path.join(ref, "ext"), os.path.join(ROOT, "legion-1_b200", "shims"), env.get("PYTHONPATH", "")]))
            tenv.pop("MASTER_ADDR", None); tenv.pop("MASTER_PORT", None)
            t1 = time.perf_counter()
            out = subprocess.run([sys.executable, os.path.join(ref, "trainers", a.trainer + ".bin"), "--class_num", str(cfg["n_class"]),
                                  "--features_num", str(D), "--train_batch_size", str(B), "--hidden_dim", "256", "--epoch", str(a.epochs),
                                  "--gpu_num", str(a.gpus)], env=tenv, cwd=a.dir, capture_output=True, text=True, timeout=1800)
            wall = time.perf_counter() - t1
            if out.returncode != 0:
                raise RuntimeError("trainer failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
            srv.wait(timeout=120)
            import re
            costs = [float(x) for x in re.findall(r"Epoch:\d+, Cost:([0-9.eE+-]+) s", out.stdout)]
            steps = [int(x) for x in re.findall(r"Train Steps: (\d+)", open(os.path.join(a.dir, "server.log")).read())]
            print(json.dumps({"what": "legion server binary -> the reference's unchanged %s.py through its unmodified ipc_service extension" % a.trainer,
                              "n_gpus": a.gpus, "workload": f"{a.config} shape, {N} nodes, {n_edges} edges, {D}-d, batch {B}/GPU, fanout {cfg['fanout']}, agg mode {a.agg_mode}, rng {a.rng}",
                              "graphsage_epoch_s": costs, "train_steps_per_epoch": steps[0] if steps else None, "trainer_wall_s": wall,
                              "model": "the trainer's own: 2 x SAGEConv/GraphConv hidden 256 (DGL stand-in), Adam, DDP over NCCL",
                              "trainer_stdout_tail": out.stdout[-400:], "server_load_presample_cache_s": t_ready}))
            shutil.rmtree(a.dir, ignore_errors=True)
            return
        cl = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--client", str(d), "--dim", str(D), "--warmup", str(a.warmup)],
                               env=env, stdout=subprocess.PIPE, text=True) for d in range(a.gpus)]
        res = []
        for p in cl:
            out, _ = p.communicate(timeout=1800)
            if p.returncode != 0:
                raise RuntimeError("client failed")
            res.append(json.loads(out.strip().splitlines()[-1]))
        srv.wait(timeout=120)
    finally:
        if srv.poll() is None:
            srv.kill()
        log.close()
    secs = max(r["seconds"] for r in res)
    batches, edges, rows = (sum(r[k] for r in res) for k in ("batches", "edges", "rows"))
    print(json.dumps({"what": "legion server binary -> ipc_service consumers (wire format of the reference)", "n_gpus": a.gpus,
                      "workload": f"{a.config} shape, {N} nodes, {n_edges} edges, {D}-d, batch {B}, fanout {cfg['fanout']}, agg mode {a.agg_mode}",
                      "batches_per_s": batches / secs, "edges_per_s": edges / secs, "feature_GBps": rows * D * 4 / secs / 1e9,
                      "ms_per_batch_per_gpu": 1e3 * secs / max(1, batches / a.gpus), "dataset_write_s": t_gen, "server_load_presample_cache_s": t_ready,
                      "server_log_tail": open(os.path.join(a.dir, "server.log")).read()[-600:]}))
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
