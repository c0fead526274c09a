#!/usr/bin/env python
"""bench.py -- Legion mini-batch hot path on N B200s (one process per GPU).

A "step" is one mini-batch of the pipeline on every rank: batch generation, k-hop neighbour sampling, dedup /
relabel and feature extraction from the NVLink-clique cache.  Headline metric (BASELINE.json): sampled edges/s
(whole job), with feature-extract GB/s and GraphSAGE epoch seconds in `extra`.

Workload at every N: BASELINE.json configs[2], the ogbn-papers100M-shaped synthetic graph (111,059,956 nodes,
~1.6 G edges, 128-d features; it fits one B200), batch 8000 per GPU, fanout [25,10] -- weak scaling: each rank
draws its own `tid % N` seed partition.  Feature cache over the N GPUs with the launcher's per-GPU budget
(legion_server.py: 38 GB): `value` is measured with the default placement (hottest rows replicated, the rest
partitioned round-robin over the clique), `extra.sharded` with the reference's pure round-robin partition
(GPUCache.cu:103-108), where (N-1)/N of the rows cross NVLink as in-kernel P2P loads.  The timed epoch is not the
presampled one (Philox counter word 1 = epoch), and before any timing every rank checks one mini-batch bit for bit
against the CPU oracle through the real peer mappings (`parity_checked`).

  python bench.py --gpus 1 --steps 50 --warmup 10
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # CPU arm: the oracle port of the reference path on all host cores
  python bench.py --config C2             # ogbn-products shape (BASELINE.json configs[1])
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE_NAME = {"C1": "100K-node", "C2": "ogbn-products", "C3": "ogbn-papers100M", "C4": "UK-Union", "C5": "Friendster"}
METRIC = "sampled edges/s"
UNIT = "edges/s"
NVLINK_NOMINAL = 900.0      # GB/s per direction per GPU (NVLink 5)

_real_stdout = None


def emit(line):
    out = _real_stdout or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(msg):
    sys.stderr.write(msg + "\n")
    sys.stderr.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--rng", default="philox", choices=["philox", "minstd"])
    ap.add_argument("--cache-frac", type=float, default=1.0, help="fraction of rows that may be cached in HBM at all (rest: host UVA tier)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nodes", type=int, default=0, help="override node count (debug)")
    ap.add_argument("--batch", type=int, default=0, help="override the seeds per mini-batch (debug: fixed per-launch cost vs work)")
    ap.add_argument("--placement", default="hybrid", choices=["sharded", "hybrid", "replicated"],
                    help="feature cache over the N GPUs: reference round-robin partition | hot rows replicated + rest partitioned | all replicated")
    ap.add_argument("--gpu-cache-gb", type=float, default=0.0,
                    help="per-GPU feature-cache budget; 0 = the launcher's 38 GB when N > 1 (legion_server.py), the whole table when N = 1")
    ap.add_argument("--topology", default="replicated", choices=["replicated", "sharded"],
                    help="CSR: a full copy in every GPU's HBM (default; 7 %% of a B200 for papers100M) | the reference's partition by topology hotness "
                         "over the clique (GPU_Memory_Graph_Storage.cu:14-133), sampled through NVLink P2P loads")
    ap.add_argument("--slot-map", default="compact", choices=["compact", "int32"],
                    help="row lookup of the feature cache: compact L2-resident placement map, rows of a class in node-id order "
                         "(lgn_place_compact; a fully resident table is addressed directly) | int32 slot_of[N] in hotness order (the reference's arrangement)")
    ap.add_argument("--no-extra-sharded", action="store_true", help="skip the second measurement with the reference partition (N > 1)")
    ap.add_argument("--no-train-epoch", action="store_true", help="skip the GraphSAGE epoch-time leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing oracle check (debug)")
    ap.add_argument("--probe", action="store_true", help="debug: time sampling-only and gather-only loops")
    ap.add_argument("--lanes", type=int, default=4, help="mini-batches in flight per GPU (batch slots)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  nvidia-smi needs ~100 ms to
    deliver its first line, so it is started early and every line is stamped on arrival; only the lines that arrived
    between begin() and end() -- the timed passes -- are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.proc, self.lines = dev, None, []
        self.t0 = self.t1 = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        self.t.join(timeout=2)
        t0, t1 = self.t0 or 0.0, (self.t1 or time.perf_counter()) + 0.03      # a line describes the 20 ms before it arrived
        inside = [ln for t, ln in self.lines if t0 <= t <= t1]
        note = None
        if not inside and self.lines:       # region shorter than the sampling period: the line nearest to it
            near = min(self.lines, key=lambda x: min(abs(x[0] - t0), abs(x[0] - t1)))
            inside, note = [near[1]], "timed region shorter than the 20 ms sampling period: nearest sample"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def workload(args):
    import legion_b200 as L
    cfg = dict(L.synth.CONFIGS[args.config])
    if args.nodes:
        cfg["n_nodes"] = args.nodes
    if args.batch:
        cfg["batch"] = args.batch
    return cfg


def workload_string(args, cfg, n_edges):
    """identical in both arms (the driver compares it)."""
    return (f"{args.config} {SHAPE_NAME.get(args.config, 'synthetic')}-shaped synthetic ({cfg['n_nodes']} nodes, {n_edges} edges, "
            f"{cfg['dim']}-d), GraphSAGE fanout {cfg['fanout']}, batch {cfg['batch']}/GPU, rng {args.rng}, sampling + feature extraction")


def _chunked_to_host(t, D):
    """device tensor [N, D] -> numpy, in 256 MB pieces (no second full-size staging copy)."""
    n = t.shape[0]
    out = np.empty((n, D), np.float32)
    rows = max(1, (1 << 28) // (4 * D))
    for lo in range(0, n, rows):
        out[lo:lo + rows] = t[lo:lo + rows].cpu().numpy()
    return out


def host_dataset(cfg, want_features=True):
    """the synthetic dataset as numpy arrays; built on the GPU when there is one (the numpy build of 1.6 G edges takes
    minutes), bit-identical either way (tests/test_oracle.py::test_synth_numpy_equals_torch)."""
    import legion_b200 as L
    N, D = cfg["n_nodes"], cfg["dim"]
    try:
        import torch
        cuda = torch.cuda.is_available()
    except Exception:      # noqa: BLE001
        cuda = False
    if not cuda or N < 5_000_000:
        return L.synth.make_dataset(N, cfg["avg_deg"], D, n_class=cfg["n_class"], with_features=want_features)
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    dmin = L.synth.calibrate_dmin(cfg["avg_deg"], N)
    ds = L.synth.make_dataset(N, cfg["avg_deg"], D, n_class=cfg["n_class"], backend="torch", device=dev, dmin_fp=dmin,
                              with_features=want_features)
    hd = L.synth.Dataset(n_nodes=N, n_edges=ds.n_edges, dim=D, n_class=cfg["n_class"], indptr=ds.indptr.cpu().numpy(),
                         indices=ds.indices.cpu().numpy(), labels=ds.labels.cpu().numpy(), train_ids=ds.train_ids.cpu().numpy(),
                         features=_chunked_to_host(ds.features, D) if want_features else None)
    del ds
    torch.cuda.empty_cache()
    return hd


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, all host threads, bounded sample
# ------------------------------------------------------------------------------------------
_cpu_state = {}


def cpu_leg(ds, cfg, seconds, steps_cap, rng_mode, seed, first_step=0):
    """returns (edges/s, GB/s, n_batches, cores, seconds).  Same semantics as the GPU path; gathers rows by
    memcpy from the host feature matrix (the reference has no CPU path of its own: SURVEY 8d)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    B, fanout = cfg["batch"], cfg["fanout"]
    key = (id(ds), rng_mode, seed)
    if key not in _cpu_state:       # sampler scratch + output buffer are allocated (and first-touched) once
        cap = O.capacity_for(B, fanout)
        out = np.zeros((cap, ds.dim), np.float32)
        _cpu_state[key] = (O.Sampler(ds.indptr, ds.indices, fanout, rng_mode=rng_mode, rng_seed=seed, n_threads=cores), out)
    smp, out = _cpu_state[key]
    train = ds.train_ids
    epoch_steps = max(1, (len(train) - 1) // B)
    edges = rows = done = 0
    t0 = time.perf_counter()
    for it in range(steps_cap):          # cycles over the epoch's batches until the time budget is spent
        step = (first_step + it) % epoch_steps
        seeds = train[step * B:(step + 1) * B]
        o = smp.sample(seeds, step=step, epoch=1)
        total = int(o["nc"][0])
        O.gather(o["sampled_ids"], 0, total, None, 1, [], ds.features, out, n_threads=cores)
        edges += int(o["ec"][0]); rows += total; done += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return edges / dt, rows * ds.dim * 4 / dt / 1e9, done, cores, dt


def reference_gpu_leg(n_batches=20):
    """The reference's OWN kernels (Kernels.cu, GPUCache.cu, ... compiled unmodified for sm_100a into oracle/_ref by
    oracle/ref_harness) driven through its own operator sequence on this GPU: presampling epoch, CandidateSelection /
    CostModel / FillUp, then n steady-state batches.  Informational: 'the kernel to beat' of BASELINE.md section 3.
    Always on the C2 (ogbn-products) shape: the harness keeps the reference's pinned-host dataset and sweeps four cache
    budgets, which would take minutes at papers100M size."""
    import legion_b200 as L
    so = os.path.join(ROOT, "oracle", "_ref", "libref_legion.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libref_legion.so not built"}
    try:
        import torch
        if not torch.cuda.is_available():
            return {"unavailable": "no CUDA device"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": repr(e)[:100]}
    cfg = dict(L.synth.CONFIGS["C2"])
    ds = L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"])
    lib = C.CDLL(so)
    lib.ref_create.restype = C.c_void_p
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    B, (f1, f2) = cfg["batch"], cfg["fanout"][:2]
    train = np.ascontiguousarray(ds.train_ids, np.int32)
    labels = np.ascontiguousarray(ds.labels[train], np.int32)
    feat_bytes, topo_bytes = ds.n_nodes * ds.dim * 4, 8 * ds.n_nodes + 4 * ds.n_edges
    steps = (len(train) - 1) // B
    cap = B * (1 + f1 + f1 * f2)
    best = None
    # The reference's CostModel zeroes a tier's gain once that tier fits completely (GPUCache.cu:744-751), so with a
    # generous budget it caches only ONE of the two tiers.  Sweep budgets and keep its fastest configuration.
    for budget in (0.85 * feat_bytes, 0.95 * feat_bytes, 1.0 * feat_bytes + 0.5 * topo_bytes, 1.02 * (feat_bytes + topo_bytes)):
        h = C.c_void_p(lib.ref_create(p(ds.indptr), p(ds.indices), C.c_int32(ds.n_nodes), C.c_int64(ds.n_edges), p(ds.features),
                                      C.c_int32(ds.dim), p(train), p(labels), C.c_int32(len(train)), C.c_int32(B), C.c_int32(f1),
                                      C.c_int32(f2), C.c_int64(int(budget))))
        ids, s_ids, d_ids = (np.zeros(cap, np.int32) for _ in range(3))
        nc, ec = np.zeros(16, np.int32), np.zeros(16, np.int32)
        trans = 0
        t0 = time.perf_counter()
        for it in range(steps):
            lib.ref_presample_batch(h, C.c_int32(it), p(ids), p(s_ids), p(d_ids), p(nc), p(ec))
            trans += int(nc[4]) + int(ec[3]) + int(ec[4])      # one UVA read transaction per indptr pair and per neighbour id
        t_pre = time.perf_counter() - t0
        qf, qt = np.zeros(ds.n_nodes, np.int32), np.zeros(ds.n_nodes, np.int32)
        ncap, ecap = C.c_int32(), C.c_int32()
        lib.ref_plan(h, C.c_uint64(trans), p(qf), p(qt), C.byref(ncap), C.byref(ecap), None, C.c_int64(0))
        ms, edges, rows = C.c_double(), C.c_int64(), C.c_int64()
        lib.ref_time_batches(h, C.c_int32(0), C.c_int32(4), C.byref(ms), C.byref(edges), C.byref(rows))          # warm-up
        lib.ref_time_batches(h, C.c_int32(4), C.c_int32(n_batches), C.byref(ms), C.byref(edges), C.byref(rows))
        res = {"ms_per_step": ms.value / n_batches, "value": edges.value / (ms.value / 1e3), "unit": UNIT,
               "feature_extract_GBps": rows.value * ds.dim * 4 / (ms.value / 1e3) / 1e9, "batches": n_batches,
               "cache_budget_bytes": int(budget), "feature_rows_cached": ncap.value, "topology_nodes_cached": ecap.value,
               "presampling_epoch_s": t_pre}
        if best is None or res["ms_per_step"] < best["ms_per_step"]:
            best = res
    best["workload"] = "C2 ogbn-products-shaped synthetic, batch 8000, fanout [25, 10] (1 GPU)"
    best["note"] = ("reference kernels recompiled unmodified for sm_100a (upstream targets sm_80), its own operator sequence on its two "
                    "event-chained streams (Server.cu:301-328) and its cache planner, minstd stream; fastest of 4 cache budgets (its cost "
                    "model caches a single tier once everything fits)")
    return best


def run_reference(args):
    """--impl reference: the reference's own semantics on the host cores (oracle port; the reference has no CPU path and its
    GPU server needs MSR access for Intel PCM).  Each step = one mini-batch (a bounded sample of the epoch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cfg = workload(args)
    ds = host_dataset(cfg)
    mode = O.RNG_PHILOX if args.rng == "philox" else O.RNG_MINSTD
    K, W = max(1, args.steps), max(0, args.warmup)
    for i in range(W):
        cpu_leg(ds, cfg, 1e9, 1, mode, 42, first_step=i)
    t_edges, t_time, gbs = 0.0, 0.0, []
    for i in range(K):
        eps, gb, n, cores, dt = cpu_leg(ds, cfg, 1e9, 1, mode, 42, first_step=W + i)
        t_edges += eps * dt; t_time += dt; gbs.append(gb)
    value = t_edges / t_time
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * t_time / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 ids / f32 rows (bit copy)", "data": "synthetic",
            "config": {"workload": workload_string(args, cfg, ds.n_edges),
                       "arm": "CPU sampling+gather, oracle port of the reference path, OpenMP on all host cores; one step = one mini-batch of one GPU's seed stream"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{K} steps x 1 batch of {cfg['batch']} seeds"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "extra": {"feature_extract_GBps": float(np.mean(gbs))}}
    del ds
    _cpu_state.clear()
    try:
        line["reference_gpu"] = reference_gpu_leg()
    except Exception as e:      # noqa: BLE001
        line["reference_gpu"] = {"unavailable": repr(e)[:200]}
    emit(line)


# ------------------------------------------------------------------------------------------
# GraphSAGE epoch time (third part of the BASELINE metric): the reference trainer's structure
# (legion_graphsage.py:36-89) fed by the pipeline, one process per GPU, DDP over NCCL
# ------------------------------------------------------------------------------------------
def graphsage_epoch(L, r, dist, world, dev, lp, NL, B, D, n_class, n_hops, train_steps, epochs=2, max_steps=400):
    import torch
    from legion_b200 import trainer

    class _V:
        def __init__(s, ptr, shape, ts):
            s.__cuda_array_interface__ = {"shape": shape, "typestr": ts, "data": (int(ptr), False), "version": 2}

    steps = min(train_steps, max_steps)
    views = []
    for q in range(NL):          # zero-copy tensors over the lane's output buffers (what ipc_service.get_next hands out)
        v = r.view(q)
        cap, rows = int(v.capacity), int(v.max_rows)
        views.append(dict(feat=torch.as_tensor(_V(v.features, (rows, D), "<f4"), device=dev),
                          lab=torch.as_tensor(_V(v.labels, (B,), "<i4"), device=dev),
                          src=torch.as_tensor(_V(v.agg_src, (cap,), "<i4"), device=dev),
                          dst=torch.as_tensor(_V(v.agg_dst, (cap,), "<i4"), device=dev)))
    torch.manual_seed(0)
    model = trainer.SAGE(D, 256, n_class, n_hops, dropout=0.5).to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index])
    opt = torch.optim.Adam(model.parameters(), lr=0.003)
    model.train()
    done_ev = [None] * NL
    times, first, last = [], None, None
    for ep in range(epochs):
        r.set_epoch(2 + ep)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(steps + NL - 1):
            if i < steps:
                q = i % NL
                if done_ev[q] is not None:
                    done_ev[q].synchronize()          # the trainer has finished reading this slot (ipc_service.synchronize())
                r.batch_generate(L.MODE_TRAIN, B, i, stream=lp[q], pipe=q)
                r.run_batch(with_features=True, stream=lp[q])
            j = i - (NL - 1)
            if j >= 0:
                q = j % NL
                nc, ec = r.read_counters(stream=lp[q], pipe=q)     # get_next(): waits for the slot, reads the counter blocks
                V = views[q]
                coo, sizes = [], []
                for layer in range(n_hops):
                    h = n_hops - 1 - layer
                    e = int(ec[3 + h])
                    coo.append((V["src"][:e], V["dst"][:e]))
                    sizes.append((int(nc[7 + 2 * h]), int(nc[5 + 2 * h])))
                total = sizes[0][0] if sizes else int(nc[4])
                loss = trainer.train_step(model, opt, V["feat"][:total], V["lab"][:int(nc[4])], coo, sizes)
                done_ev[q] = torch.cuda.Event()
                done_ev[q].record()
                if first is None:
                    first = float(loss)
                last = loss
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    scale = train_steps / steps
    return {"graphsage_epoch_s": times[-1] * scale, "graphsage_first_epoch_s": times[0] * scale, "graphsage_steps_per_epoch": train_steps,
            "graphsage_steps_timed": steps,
            "graphsage_model": f"{n_hops}-layer SAGEConv(mean) hidden 256, Adam, fp32, {'DDP' if world > 1 else 'single GPU'}; in-process consumer of the "
                               "batch slots (tools/server_e2e.py times the server binary + trainer processes over the IPC wire)",
            "graphsage_loss_first": first, "graphsage_loss_last": float(last)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Cache:
    """one placement of the feature cache over the clique, bound to the runner."""

    def __init__(self):
        self.shards, self.imported, self.my_shard, self.slot_of = [], [], None, None
        self.cap = self.n_repl = self.kg_bind = 0
        self.placement = ""
        self.vmm = False


def run_b200(args):
    import torch
    import torch.distributed as dist
    import legion_b200 as L
    from legion_b200 import cluster

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # LGN_BENCH_BACKEND=gloo: diagnostic run without any NCCL communicator in the process; the GraphSAGE/DDP leg
        # needs NCCL and is skipped then
        if os.environ.get("LGN_BENCH_BACKEND", "nccl") == "gloo":
            dist.init_process_group("gloo")
            args.no_train_epoch = True
        else:
            dist.init_process_group("nccl", device_id=dev)
    rdev = dev if world == 1 or dist.get_backend() == "nccl" else torch.device("cpu")     # where reduced scalars live
    cfg = workload(args)
    N, D, B, fanout = cfg["n_nodes"], cfg["dim"], cfg["batch"], cfg["fanout"]
    rng_mode = L.RNG_PHILOX if args.rng == "philox" else L.RNG_MINSTD
    row_bytes = 4 * D
    t_start = time.perf_counter()

    # ---- dataset, resident in HBM before the timed region --------------------------------
    dmin = L.synth.calibrate_dmin(cfg["avg_deg"], N)
    ds = L.synth.make_dataset(N, cfg["avg_deg"], D, n_class=cfg["n_class"], backend="torch", device=dev, dmin_fp=dmin)
    torch.cuda.synchronize()
    t_dataset = time.perf_counter() - t_start
    my_train = cluster.partition_seeds(ds.train_ids, world, rank).contiguous()        # GPUGraphStore.cu:338
    my_labels = ds.labels[my_train.long()].contiguous()
    train_steps = cluster.train_steps(my_train.numel(), B, dist if world > 1 else None)    # CUDA_IPC_Service.cu:88

    r = L.Runner(N, D, B, fanout, device=local, part=rank, rng_mode=rng_mode, rng_seed=42, enable_hotness=True, n_lanes=args.lanes)
    topo = (ds.indptr, ds.indices)
    if os.environ.get("LGN_BENCH_TOPO_ALLOC") == "vmm":     # experiment: CSR in VMM allocations of 512 MB granules (address-translation reach)
        topo = []
        for t, dt in ((ds.indptr, np.int64), (ds.indices, np.int32)):
            a, fd_, _mb = L.shared_alloc((t.numel(),), dt)
            os.close(fd_)
            L._lib.check(L.lib().lgn_copy_d2d(C.c_void_p(a.ptr), C.c_void_p(t.data_ptr()), C.c_int64(t.numel() * t.element_size())), "copy CSR")
            topo.append(a)
    r.bind_topology(topo[0], topo[1])            # topology replicated in HBM (7 % of one B200 even for papers100M)
    r.bind_seeds(L.MODE_TRAIN, my_train, my_labels)
    NL = args.lanes
    lanes = [torch.cuda.Stream(device=dev) for _ in range(NL)]     # one sampling stream per batch slot
    lp = [x.cuda_stream for x in lanes]
    stream, sp = lanes[0], lp[0]
    stream2 = torch.cuda.Stream(device=dev)       # consumer-side stream of the e2e leg (result read-back)
    sp2 = stream2.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- presampling epoch (epoch 0) -> hotness -> all-reduce -> hot order -----------------
    r.set_epoch(0)
    t_pre = time.perf_counter()
    presample_steps = min(train_steps, int(os.environ.get("LGN_BENCH_PRESAMPLE_STEPS", train_steps)))   # profiling runs shorten the epoch
    for step in range(presample_steps):
        r.batch_generate(L.MODE_TRAIN, B, step, stream=lp[step % NL], pipe=step % NL)
        r.run_batch(with_features=False, is_presc=True, stream=lp[step % NL])
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t_pre
    r.set_dedup_capacity(max(1, r.max_ids()))     # hash dedup table: 2.5 x the largest presampled batch (no-op for the direct map)
    nh, _th = r.hotness()
    if world > 1:   # the path's one collective: all-reduce of the hotness histogram (replaces aggregate_access, GPUCache.cu:624-647)
        how = os.environ.get("LGN_BENCH_HOTNESS", "nccl" if dist.get_backend() == "nccl" else "host")
        nh_t = torch.as_tensor(cluster._DeviceView(nh.ptr, N), device=dev)
        if how == "none":       # diagnostic: no reduction at all, identical (id-ordered) placement on every rank
            nh_t.zero_()
        elif how == "host":     # diagnostic: reduce through host memory (gloo needs CPU tensors)
            h = nh_t.cpu()
            grp = dist.new_group(backend="gloo") if dist.get_backend() == "nccl" else None
            dist.all_reduce(h, group=grp)
            nh_t.copy_(h)
        elif how == "torch":    # the same reduction through torch.distributed's communicator
            cluster.allreduce_hotness(dist, nh, n=N, device=dev)
        else:                   # default: ncclAllReduce inside liblegion_b200.so (lgn_comm_*), torch only ships the unique id
            cluster.native_allreduce_u32(dist, nh.ptr, N, stream=sp)
        torch.cuda.synchronize()
    order, hot_sorted = L.hot_order(nh, want_sorted=True)
    kg = world
    topo_note = "topology replicated in HBM"
    topo_keep = []
    if args.topology == "sharded" and world > 1:
        # the reference's topology cache: nodes in topology-hotness order dealt round-robin over the clique, GPU j holds the
        # adjacency lists of ranks j, j + Kg, ..; the sampler resolves a frontier node through slot_of and reads the indptr
        # pair and the neighbour ids from the owner's shard (local HBM or NVLink)
        cluster.native_allreduce_u32(dist, _th.ptr, N, stream=sp)
        torch.cuda.synchronize()
        order_t = L.hot_order(_th)
        cap_t = cluster.capacity_for(N, kg)
        ip_s, ix_s, n_ix = L.fill_topo_shard(order_t, cap_t, kg, rank, ds.indptr, ds.indices)
        torch.cuda.synchronize()
        tabs = []
        for arr, shape, dt in ((ip_s, (cap_t + 1,), np.int64), (ix_s, None, np.int32)):
            h = (C.c_uint8 * 64)()
            L._lib.check(L.lib().lgn_ipc_export(C.c_void_p(arr.ptr), h), "ipc_export")
            meta = [None] * world
            dist.all_gather_object(meta, (bytes(h), int(arr.shape[0])))
            lst = []
            for j in range(world):
                if j == rank:
                    lst.append(arr)
                    continue
                pp = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(meta[j][0])
                L._lib.check(L.lib().lgn_ipc_import(hb, C.byref(pp)), "ipc_import")
                lst.append(L.DevArray((meta[j][1],), dt, ptr=pp.value, owner=False))
            tabs.append(lst)
        slot_t = L.place(order_t, cap_t, kg)
        r.bind_topology_cache(tabs[0], tabs[1], slot_t, cap_t)
        topo_keep = [ip_s, ix_s, slot_t, tabs]
        topo_note = f"topology partitioned over {kg} GPUs by topology hotness ({n_ix} neighbour ids in this GPU's shard), sampled over NVLink"
    n_cached = int(N * args.cache_frac)
    budget_gb = args.gpu_cache_gb if args.gpu_cache_gb > 0 else (38.0 if world > 1 else N * row_bytes / 1e9 + 1.0)
    budget_rows = int(budget_gb * 1e9 // row_bytes)

    hbm_peak, peak_src = peaks()

    # ---- link probes for the hit-mix roofline (measured, not nominal) ----
    def copy_GBps(dst, src, nbytes, reps=4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L._lib.check(L.lib().lgn_copy_async(C.c_void_p(dst), C.c_void_p(src), C.c_int64(nbytes), C.c_void_p(sp)), "copy")
        e0.record(stream)
        for _ in range(reps):
            L._lib.check(L.lib().lgn_copy_async(C.c_void_p(dst), C.c_void_p(src), C.c_int64(nbytes), C.c_void_p(sp)), "copy")
        e1.record(stream)
        e1.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) / 1e3) / 1e9

    probe_bytes = 256 << 20
    pin = L.MappedHostArray((probe_bytes // 4,), np.float32)
    scratch = L.DevArray((probe_bytes // 4,), np.float32)
    pin.array[:] = 1.0
    pcie_h2d = copy_GBps(scratch.ptr, pin.host_ptr, probe_bytes)
    pcie_d2h = copy_GBps(pin.host_ptr, scratch.ptr, probe_bytes)
    pin.free()

    # ---- cache placements ------------------------------------------------------------------
    host_tier = [None]

    def build_cache(placement):
        c = Cache()
        c.placement = placement
        if placement == "replicated":
            c.n_repl = min(n_cached, budget_rows)
            c.cap = max(1, c.n_repl)
        elif placement == "hybrid" and kg > 1:
            # the B200 placement model (lgn_plan_hybrid): replicated / partitioned / host split that
            # minimises the expected gather time of the presampled epoch under the three tier rates (payload GB/s per GPU:
            # HBM copy / 2, the all-to-all random-row NVLink rate tools/peer_probe measured, the PCIe rate measured above)
            c.n_repl, c.cap, _cost = L.plan_hybrid(hot_sorted, D, int(min(budget_gb * 1e9, n_cached * row_bytes + row_bytes)), kg,
                                                   hbm_peak / 2, 640.0, pcie_h2d, stream=sp)
            c.n_repl = min(c.n_repl, n_cached)                   # --cache-frac: rows beyond n_cached stay on the host tier
            if budget_rows * kg >= n_cached:                     # the clique can hold every cacheable row: never spill to the host
                c.n_repl = min(c.n_repl, max(0, (budget_rows * kg - n_cached) // (kg - 1)))     # tier for the sake of one more replica
                c.cap = c.n_repl + cluster.capacity_for(n_cached - c.n_repl, kg)
            c.cap = min(c.cap, c.n_repl + cluster.capacity_for(n_cached - c.n_repl, kg)) if n_cached > c.n_repl else max(1, c.n_repl)
        else:                            # reference partition (GPUCache.cu:103-108); one GPU: everything the budget allows, locally
            c.n_repl = 0
            part_rows = min(n_cached, budget_rows * kg)
            c.cap = cluster.capacity_for(part_rows, kg)
        if kg > 1 and c.n_repl >= n_cached:      # nothing is partitioned: every GPU holds the whole cached set, no peer shards to bind
            c.kg_bind, part_bind = 1, 0
        else:
            c.kg_bind, part_bind = kg, rank
        r.set_part(part_bind)
        cached_rows = min(N, c.n_repl + (c.cap - c.n_repl) * c.kg_bind)
        c.compact = args.slot_map == "compact"
        c.identity = c.compact and c.kg_bind == 1 and cached_rows >= N       # whole table resident: rows addressed by node id
        if os.environ.get("LGN_BENCH_NO_IDENTITY"):                         # experiment: go through the placement map anyway
            c.identity = False
        if c.identity:
            c.slot_of = None
        elif c.compact:
            try:
                c.slot_of = L.place_compact(order, *cluster.compact_split(N, c.n_repl, c.cap, c.kg_bind), stream=sp)
            except L.LegionError as e:       # argument validation only (identical on every rank): fall back to the int32 table, say so
                log(f"rank {rank}: compact placement refused ({e}); using the int32 slot table")
                c.compact = False
        if not c.identity and not c.compact:
            c.slot_of = L.place_hybrid(order, c.cap, c.kg_bind, c.n_repl, part_bind)
        base = ds.features
        if cached_rows < N:       # misses: pinned host memory over UVA (only then is the 4*N*D-byte host copy made)
            if host_tier[0] is None:
                try:
                    import psutil
                    avail = psutil.virtual_memory().available
                except Exception:      # noqa: BLE001
                    avail = 1 << 62
                if N * row_bytes * world > 0.6 * avail:      # every rank pins its own host copy of the feature matrix
                    raise SystemExit(f"host tier needs {N * row_bytes * world / 1e9:.0f} GB of pinned host memory over {world} ranks, "
                                     f"{avail / 1e9:.0f} GB available: raise --gpu-cache-gb / --cache-frac or lower --nodes")
                host_tier[0] = L.MappedHostArray((N, D), np.float32)
                rows = max(1, (1 << 28) // row_bytes)
                for lo in range(0, N, rows):
                    host_tier[0].array[lo:lo + rows] = ds.features[lo:lo + rows].cpu().numpy()
            base = host_tier[0]
        r.bind_features(base)
        if c.identity:
            c.my_shard = None
            c.shards = [ds.features]
            c.cap = c.n_repl = N
            r.bind_feature_cache_compact(c.shards, None, N, N)
            barrier()
            return c
        c.vmm = os.environ.get("LGN_BENCH_SHARD_ALLOC", "ipc") == "vmm"      # shard through the VMM API (512 MB granules) instead of cudaMalloc
        shard_fd = mapped_bytes = None
        out = None
        if c.vmm:
            out, shard_fd, mapped_bytes = L.shared_alloc((c.cap, D), np.float32)
        if c.compact:
            c.my_shard = L.fill_feature_shard_compact(c.slot_of, N, c.n_repl, c.kg_bind, part_bind, ds.features, D, c.cap, out=out)
        else:
            c.my_shard = L.fill_feature_shard_hybrid(order, c.cap, c.kg_bind, part_bind, c.n_repl, ds.features, D, out=out)
        c.shards = [c.my_shard]
        if c.kg_bind == 1 and c.vmm:
            os.close(shard_fd)
        if c.kg_bind > 1 and c.vmm:
            torch.cuda.synchronize()
            fds = cluster.exchange_fds(dist, shard_fd)
            sizes = [None] * world
            dist.all_gather_object(sizes, int(mapped_bytes))
            c.shards = []
            for j in range(world):
                if j == rank:
                    c.shards.append(c.my_shard)
                else:
                    a = L.shared_import(fds[j], sizes[j], (c.cap, D), np.float32)
                    c.imported.append(a)
                    c.shards.append(a)
                    os.close(fds[j])
            os.close(shard_fd)
        elif c.kg_bind > 1:   # peer shards: CUDA IPC handles exchanged once, then plain P2P loads inside the gather kernel
            torch.cuda.synchronize()
            h = (C.c_uint8 * 64)()
            L._lib.check(L.lib().lgn_ipc_export(C.c_void_p(c.my_shard.ptr), h), "ipc_export")
            handles = cluster.exchange_handles(dist, bytes(h))
            c.shards = []
            for j in range(world):
                if j == rank:
                    c.shards.append(c.my_shard)
                    continue
                p = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(handles[j])
                L._lib.check(L.lib().lgn_ipc_import(hb, C.byref(p)), "ipc_import")
                c.imported.append(p)
                c.shards.append(L.DevArray((c.cap, D), np.float32, ptr=p.value, owner=False))
        if c.compact:
            r.bind_feature_cache_compact(c.shards, c.slot_of, c.n_repl, c.cap)
        else:
            r.bind_feature_cache(c.shards, c.slot_of, c.cap)
        barrier()
        return c

    def drop_cache(c):
        barrier()
        for p in c.imported:
            if c.vmm:
                L.shared_free(p)
            else:
                L.lib().lgn_ipc_close(p)
        barrier()                      # every importer has unmapped before any owner frees
        if c.my_shard is not None:
            if c.vmm:
                L.shared_free(c.my_shard)
            else:
                c.my_shard.free()
        if c.slot_of is not None:
            c.slot_of.free()
        c.shards, c.imported = [], []

    # ---- parity: one mini-batch of THIS rank against the CPU oracle, through the real tier mappings ----
    oracle_state = {}
    shm_files = []

    def host_csr():
        """the CSR in host memory for the oracle; with several ranks ONE copy in /dev/shm is shared by all of them
        (8 private copies of papers100M's 7.4 GB CSR would not fit every box)."""
        if world == 1:
            return ds.indptr.cpu().numpy(), ds.indices.cpu().numpy()
        need = 8 * (N + 1) + 4 * ds.n_edges
        where = None
        if rank == 0:      # a tmpfs that is too small would kill the writer with SIGBUS: look before writing
            import shutil
            import tempfile
            for cand in ("/dev/shm", tempfile.gettempdir()):
                try:
                    if shutil.disk_usage(cand).free > 1.1 * need:
                        where = cand
                        break
                except OSError:
                    pass
        tok = [(os.urandom(4).hex(), where) if rank == 0 else None]
        dist.broadcast_object_list(tok, src=0)
        tok, where = tok[0]
        if where is None:       # no shared place large enough: private copies
            return ds.indptr.cpu().numpy(), ds.indices.cpu().numpy()
        tok = [tok]
        paths = ["%s/lgn_bench_%s_%s" % (where, tok[0], nm) for nm in ("indptr", "indices")]
        specs = [(ds.indptr, np.int64), (ds.indices, np.int32)]
        if rank == 0:
            for path, (t, dt) in zip(paths, specs):
                m = np.memmap(path, dtype=dt, mode="w+", shape=(t.numel(),))
                step_ = 1 << 26
                for lo in range(0, t.numel(), step_):
                    m[lo:lo + step_] = t[lo:lo + step_].cpu().numpy()
                m.flush()
                del m
                shm_files.append(path)
        dist.barrier()
        return tuple(np.memmap(path, dtype=dt, mode="r", shape=(t.numel(),)) for path, (t, dt) in zip(paths, specs))

    def parity_check(c, step=0, epoch=1):
        from oracle import oracle as O
        if "smp" not in oracle_state:
            ip_h, ix_h = host_csr()
            oracle_state["smp"] = O.Sampler(ip_h, ix_h, fanout,
                                            rng_mode=O.RNG_PHILOX if args.rng == "philox" else O.RNG_MINSTD, rng_seed=42,
                                            n_threads=max(1, (os.cpu_count() or 8) // max(1, world)))
        seeds = my_train[step * B:(step + 1) * B].cpu().numpy()
        labels = my_labels[step * B:(step + 1) * B].cpu().numpy()
        if "want" not in oracle_state:
            oracle_state["want"] = oracle_state["smp"].sample(seeds, step=step, epoch=epoch)
        want = oracle_state["want"]
        r.set_epoch(epoch)
        r.batch_from_host(seeds, labels, step=step, stream=sp, pipe=0)
        r.run_batch(with_features=True, stream=sp)
        got = r.fetch(with_features=True, stream=sp)
        total, n_e = int(want["nc"][0]), int(want["ec"][0])
        ok = np.array_equal(got["nc"], want["nc"]) and np.array_equal(got["ec"], want["ec"])
        ok = ok and np.array_equal(got["sampled_ids"], want["sampled_ids"][:total])
        for k in ("agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off"):
            ok = ok and np.array_equal(got[k], want[k][:n_e])
        ok = ok and np.array_equal(got["labels"], labels)
        if ok:       # features are closed-form (synth.features_block): every gathered row must equal its node's row, bit for bit
            ids = want["sampled_ids"][:total].astype(np.uint64)
            exp = ((ids[:, None] * np.uint64(2654435761) + np.arange(D, dtype=np.uint64)[None, :] * np.uint64(40503)) & np.uint64(0x7FFFFF)) \
                + np.uint64(0x3F000000)
            ok = np.array_equal(got["features"].view(np.uint32), exp.astype(np.uint32))
        flag = torch.tensor([1 if ok else 0], device=rdev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            raise SystemExit(f"rank {rank}: mini-batch differs from the CPU oracle (placement {c.placement}, own result {'ok' if ok else 'WRONG'}) "
                             "-- refusing to time a wrong result")
        return {"rows": total, "edges": n_e}

    # ---- measurement of one placement ------------------------------------------------------
    def step_resident(i):
        r.batch_generate(L.MODE_TRAIN, B, i % train_steps, stream=lp[i % NL], pipe=i % NL)
        r.run_batch(with_features=True, stream=lp[i % NL])

    n_pin = min(train_steps, max(64, args.steps + args.warmup + 8))
    seeds_pin = torch.empty((n_pin, B), dtype=torch.int32).pin_memory()
    labels_pin = torch.empty((n_pin, B), dtype=torch.int32).pin_memory()
    seeds_pin.copy_(my_train[:n_pin * B].view(n_pin, B).cpu())
    labels_pin.copy_(my_labels[:n_pin * B].view(n_pin, B).cpu())

    def step_e2e(i):
        # double-buffered consumer (like the reference's trainer handshake): enqueue batch i from pinned host seeds (H2D),
        # then read the oldest in-flight batch's result block (D2H + host sync)
        j = i % n_pin
        r.batch_from_host(seeds_pin[j], labels_pin[j], step=j, stream=lp[i % NL], pipe=i % NL)
        r.run_batch(with_features=True, stream=lp[i % NL])
        return r.read_counters(stream=sp2, pipe=(i + 1) % NL)

    host_ms = [0.0]

    def timed(fn, K, W, profile=False):
        for i in range(W):
            fn(i)
        barrier()
        if profile:
            r.profile_enable(K * 8 + 16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for i in range(W, W + K):
            fn(i)
        host_ms[0] = 1e3 * (time.perf_counter() - h0) / K
        for q in range(NL):              # the timed region ends when every slot's last gather is done
            r.wait_pipe(q, stream=sp)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            tl = r.profile_timeline(K * 8 + 16)          # (kind, slot, begin_ms, end_ms) of every operator of the timed region
            iv = sorted((a, b) for kind, _, a, b in tl if kind == 2)
            busy, cur_a, cur_b = 0.0, None, None
            for a, b in iv:                               # union of the gather launches' intervals
                if cur_b is None or a > cur_b:
                    busy += (cur_b - cur_a) if cur_b is not None else 0.0
                    cur_a, cur_b = a, b
                else:
                    cur_b = max(cur_b, b)
            busy += (cur_b - cur_a) if cur_b is not None else 0.0
            prof = r.profile_collect() + (busy,)
            r.profile_enable(0)
        t = torch.tensor([ms], device=rdev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), prof

    K, W = args.steps, max(args.warmup, 3)

    def count_work(K, W):
        """work done in the timed steps (deterministic: replay the same steps untimed and read the counters)."""
        edges = rows = 0
        hop_items, hop_edges, hop_new = (np.zeros(len(fanout), np.int64) for _ in range(3))
        for i in range(W, W + K):
            step_resident(i)
            nc, ec = r.read_counters(stream=sp)
            edges += int(ec[0]); rows += int(nc[0])
            prev_e = prev_items = 0
            for h in range(len(fanout)):
                e_h = int(ec[3 + h]) - prev_e
                hop_items[h] += B if h == 0 else prev_items
                hop_edges[h] += e_h; hop_new[h] += int(nc[6 + 2 * h])
                prev_items, prev_e = e_h, int(ec[3 + h])
        assert r.status(stream=sp) == 0, "device-side capacity overflow"
        return edges, rows, hop_items, hop_edges, hop_new

    def nvlink_probe(c):
        """peer copy rate of THIS box: every rank copies 256 MB out of its right neighbour's shard at the same time."""
        if c.kg_bind < 2:
            return None
        nb = min(probe_bytes, c.cap * row_bytes)
        barrier()
        g = copy_GBps(scratch.ptr, c.shards[(rank + 1) % world].ptr, nb)
        t = torch.tensor([g], device=rdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def hit_mix(tiers, rows, ms, nvl):
        tsum = max(1, sum(tiers))
        h_local, h_peer, h_host = tiers[0] / tsum, tiers[1] / tsum, tiers[2] / tsum
        inv = h_local / (hbm_peak / 2) + h_peer / nvl + h_host / pcie_h2d
        roof = 1.0 / inv if inv > 0 else hbm_peak / 2          # payload GB/s per GPU
        inv_nom = h_local / (hbm_peak / 2) + h_peer / NVLINK_NOMINAL + h_host / pcie_h2d
        payload = rows * row_bytes / (ms / 1e3) / 1e9
        return {"local": h_local, "peer": h_peer, "host": h_host, "payload_roof_GBps_per_gpu": roof,
                "achieved_payload_GBps_per_gpu": payload, "frac": payload / roof,
                "frac_vs_nominal_900": payload * inv_nom,
                "peaks_GBps": {"hbm_copy": hbm_peak, "nvlink_peer_copy_measured": nvl, "nvlink_nominal": NVLINK_NOMINAL,
                               "pcie_h2d_pinned_measured": pcie_h2d, "pcie_d2h_pinned_measured": pcie_d2h},
                "note": "achieved = feature payload of this rank's step / step time (sampling included); roof = 1 / (h_local/(HBM copy/2) + "
                        "h_peer/NVLink + h_host/PCIe) with the link rates measured on this box"}

    def measure_light(c):
        """the graph-replayed timed region only (second placement)."""
        r.set_epoch(1)
        r.tier_counts(reset=True, stream=sp)
        ms_total, _ = timed(step_resident, K, W)
        tiers = r.tier_counts(reset=True, stream=sp)
        edges, rows, *_ = count_work(K, W)
        r.tier_counts(reset=True, stream=sp)
        tot = torch.tensor([edges, rows], device=rdev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot)
        nvl = nvlink_probe(c) or 770.0
        return {"placement": c.placement, "rows_replicated": int(c.n_repl), "rows_per_shard": int(c.cap),
                "slot_map": "none" if c.identity else "compact" if c.compact else "int32",
                "value": float(tot[0].item()) / (ms_total / 1e3), "unit": UNIT, "ms_per_step": ms_total / K,
                "feature_extract_GBps": float(tot[1].item()) * row_bytes / (ms_total / 1e3) / 1e9,
                "tier_rows_per_timed_region": [int(t * K / (K + W)) for t in tiers],
                "hit_mix": hit_mix(tiers, rows, ms_total, nvl)}

    if args.probe:
        c = build_cache(args.placement)
        return probe_mode(L, r, dist, world, rank, c, lp, lanes, stream, sp, NL, B, D, train_steps, barrier, step_resident)

    # ======================= primary placement ==============================================
    clocks = ClockSampler(local)
    clocks.start()
    c = build_cache(args.placement)
    parity = None
    if not args.no_parity:
        t0 = time.perf_counter()
        parity = parity_check(c)
        parity["seconds"] = time.perf_counter() - t0
    r.set_epoch(1)                                    # the timed epoch is not the presampled one
    r.tier_counts(reset=True, stream=sp)
    clocks.begin()
    ms_prof, prof = timed(step_resident, K, W, profile=True)      # instrumented pass: per-operator CUDA events, eager launches
    host_enqueue_ms = host_ms[0]
    tiers = r.tier_counts(reset=True, stream=sp)
    ms_total, _ = timed(step_resident, K, W)          # THE timed region: the same K steps as the product runs them (CUDA-graph replay)
    host_enqueue_plain_ms = host_ms[0]
    ms_e2e, _ = timed(step_e2e, K, W)
    edges, rows, hop_items, hop_edges, hop_new = count_work(K, W)
    tot = torch.tensor([edges, rows], device=rdev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    job_edges, job_rows = float(tot[0].item()), float(tot[1].item())
    value = job_edges / (ms_total / 1e3)
    e2e_value = job_edges / (ms_e2e / 1e3)
    feat_gbps = job_rows * row_bytes / (ms_total / 1e3) / 1e9
    nvl = nvlink_probe(c) or 770.0

    # long run of the same region: DVFS / pipeline-fill effects that 50 steps cannot show
    long_K = int(os.environ.get("LGN_BENCH_LONG_STEPS", "2000"))
    ms_long = None
    if long_K > 0:
        ms_long, _ = timed(step_resident, long_K, W)
    clocks.end()                                      # samples span the three timed passes and the long run of the same region
    clk = clocks.stop()

    # ---- the dominant kernel timed ALONE (no other batch in flight) ------------------------
    alone_ms = alone_calls = alone_rows = 0
    for i in range(W, W + min(K, 16)):
        q = i % NL
        r.batch_generate(L.MODE_TRAIN, B, i % train_steps, stream=lp[q], pipe=q)
        r.run_batch(with_features=False, stream=lp[q])
        nc_, _ = r.read_counters(stream=lp[q], pipe=q)
        barrier()
        r.select_pipe(q)
        r.profile_enable(8)
        r.gather_all(stream=lp[q])           # the launches lgn_run_batch issues for the batch's feature extraction
        torch.cuda.synchronize()
        ms_k, calls_k = r.profile_collect()
        r.profile_enable(0)
        alone_ms += ms_k[2]; alone_calls += calls_k[2]; alone_rows += int(nc_[0])

    # ---- roofline of the dominant kernel (feature gather), timed live with CUDA events ----
    ms_kind, calls, gather_busy_ms = prof
    gather_ms, gather_calls = ms_kind[2], calls[2]
    alg_bytes = rows * (2 * row_bytes + 8)                   # 2r + 8 per unique row (this rank)
    achieved = alg_bytes / (gather_busy_ms / 1e3) / 1e9 if gather_busy_ms > 0 else 0.0
    achieved_sum = alg_bytes / (gather_ms / 1e3) / 1e9 if gather_ms > 0 else 0.0
    samp_bytes = float(sum(16 * hop_items[h] + 12 * hop_edges[h] + 4 * hop_new[h] for h in range(len(fanout))))
    traffic = wasted = None
    tp = os.path.join(ROOT, "profiles", "r02_gather_traffic.json")
    if os.path.exists(tp) and gather_calls:      # DRAM bytes per row from this round's ncu --set full capture of the same kernel
        tj = json.load(open(tp)).get(args.config)
        if tj:
            traffic = tj["traffic_bytes_per_row"] * rows / gather_calls
            wasted = tj["traffic_bytes_per_row"] / (2 * row_bytes + 8)
    scale = (K + W) / K      # tiers were counted over warm-up + timed steps of the instrumented pass
    roofline = {"bound": "hbm", "kernel": r.gather_kernel_name(), "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_over_algorithmic": wasted, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(1, gather_calls),
                "step": {"achieved": alg_bytes / (ms_total / 1e3) / 1e9, "frac": alg_bytes / (ms_total / 1e3) / 1e9 / hbm_peak,
                         "note": "the same algorithmic bytes over the WHOLE step time of the timed region (sampling included): what a reader "
                                 "recomputes from rows x bytes / ms_per_step"},
                "alone": {"achieved": alone_rows * (2 * row_bytes + 8) / (alone_ms / 1e3) / 1e9 if alone_ms else None,
                          "frac": alone_rows * (2 * row_bytes + 8) / (alone_ms / 1e3) / 1e9 / hbm_peak if alone_ms else None,
                          "launches": int(alone_calls), "avg_launch_us": 1e3 * alone_ms / max(1, alone_calls),
                          "note": "same kernel, same rows, same launches, replayed with nothing else in flight"},
                "note": "CUDA events around every gather launch inside the instrumented timed region; %d batches are in flight, so "
                        "launches of different batches overlap: achieved = algorithmic bytes / time with >= 1 gather launch executing "
                        "(overlap counted once), per_launch_sum = same bytes / sum of launch durations; sampling kernels of the other "
                        "batches run concurrently and share L2/HBM (alone: see `alone`)" % NL,
                "launches": int(gather_calls), "avg_launch_us": 1e3 * gather_busy_ms / max(1, gather_calls),
                "per_launch_sum": {"achieved": achieved_sum, "frac": achieved_sum / hbm_peak, "avg_launch_us": 1e3 * gather_ms / max(1, gather_calls)},
                "algorithmic_bytes_per_row": 2 * row_bytes + 8,
                "hit_mix": hit_mix(tiers, rows * scale, ms_total * scale, nvl),
                "sampler": {"ms_per_step": ms_kind[1] / K, "algorithmic_GBps": samp_bytes / (ms_kind[1] / 1e3) / 1e9 if ms_kind[1] else 0.0},
                "share_of_step": {"gather_ms": gather_ms / K, "sample_ms": ms_kind[1] / K, "begin_ms": ms_kind[0] / K,
                                  "end_ms": ms_kind[3] / K, "step_ms": ms_prof / K,
                                  "note": "operator durations from the instrumented pass over the same K steps (per-operator CUDA events need eager launches, "
                                          "so that pass is a few % slower than the graph-replayed timed region); gather and sampling overlap on separate streams"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 ids / f32 rows (bit copy)", "data": "synthetic",
            "config": {"workload": workload_string(args, cfg, ds.n_edges),
                       "parallelism": f"dp{world}: seeds tid%{world}, feature cache {c.placement} over {kg} GPU(s), {topo_note}",
                       "placement": c.placement, "rows_replicated": int(c.n_repl), "rows_per_shard": int(c.cap), "cache_frac": args.cache_frac,
                       "slot_map": ("none: whole table resident in node-id order, rows addressed directly" if c.identity else
                                    "compact L2-resident placement map (lgn_place_compact)" if c.compact else "int32 slot_of[N], hotness order"),
                       "gpu_cache_budget_GB": budget_gb, "batches_in_flight": NL,
                       "l2": "inputs larger than L2: working set (feature shard %.2f GB + %.2f GB CSR) against 126 MB; "
                             "consecutive steps touch different rows" % (c.cap * row_bytes / 1e9, (8 * N + 4 * ds.n_edges) / 1e9),
                       "global_batch": B * world, "train_steps_per_epoch": train_steps,
                       "timed_epoch": "epoch 1 (presampling saw epoch 0: other neighbourhoods)"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 128,
                    "ms_per_step": ms_e2e / K,
                    "note": "lgn_batch_from_host (pinned seeds+labels H2D) + lgn_run_batch + lgn_read_counters (D2H, sync) every step; "
                            "extracted features stay in HBM for the on-GPU trainer, as in the reference"},
            "gpu_launches": int(K * r.launches_per_batch()),
            "parity_checked": parity is not None,
            "roofline": roofline,
            "extra": {"feature_extract_GBps": feat_gbps, "unique_rows_per_step": rows / K, "edges_per_step": edges / K,
                      "graphsage_dataloading_epoch_s": train_steps * ms_total / K / 1e3,
                      "presampling_epoch_s": t_pre, "dataset_build_s": t_dataset,
                      "tier_rows_per_timed_region": [int(t / scale) for t in tiers],
                      "host_enqueue_ms_per_step": host_enqueue_ms, "host_enqueue_ms_per_step_unprofiled": host_enqueue_plain_ms,
                      "ms_per_step_instrumented_pass": ms_prof / K,
                      "parity": parity and dict(parity, what="nc, ec, sampled_ids, 4 edge arrays, labels, features of batch 0 of epoch 1 of every rank "
                                                            "== CPU oracle, bit for bit, read through the bound tier mappings")}}
    if ms_long is not None:
        line["extra"]["long_run"] = {"steps": long_K, "ms_per_step": ms_long / long_K}

    if not args.no_train_epoch:
        try:
            line["extra"].update(graphsage_epoch(L, r, dist, world, dev, lp, NL, B, D, cfg["n_class"], len(fanout), train_steps))
        except Exception as e:      # noqa: BLE001  (the model leg must never invalidate the data-path numbers)
            line["extra"]["graphsage_epoch_error"] = repr(e)[:200]

    # ======================= the reference partition, real NVLink peer reads ================
    if world > 1 and c.placement != "sharded" and not args.no_extra_sharded:
        try:
            drop_cache(c)
            c = build_cache("sharded")
            if not args.no_parity:
                parity_check(c)
            sh = measure_light(c)
            sh["parity_checked"] = not args.no_parity
            line["extra"]["sharded"] = sh
        except SystemExit:
            raise
        except Exception as e:      # noqa: BLE001
            line["extra"]["sharded"] = {"error": repr(e)[:300]}
    if world > 1:
        line["extra"]["peer_parity"] = "bit-exact" if not args.no_parity else "skipped"

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) -------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:      # noqa: BLE001
            avail = 1 << 62
        need = N * row_bytes + 8 * N + 4 * ds.n_edges
        if avail > 1.4 * need:
            smp = oracle_state.get("smp")
            hd = L.synth.Dataset(indptr=smp.indptr if smp else ds.indptr.cpu().numpy(), indices=smp.indices if smp else ds.indices.cpu().numpy(),
                                 features=_chunked_to_host(ds.features, D), train_ids=my_train.cpu().numpy(), dim=D)
            eps, gb, n, cores, dt = cpu_leg(hd, cfg, args.cpu_seconds, 100000, O.RNG_PHILOX if args.rng == "philox" else O.RNG_MINSTD, 42)
            line["cpu_baseline"] = {"value": eps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} batches of {B} seeds of the same workload in {dt:.1f} s (oracle, OpenMP)",
                                    "feature_extract_GBps": gb}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"skipped: host copy of the dataset needs {need / 1e9:.0f} GB, {avail / 1e9:.0f} GB available"}
    line["extra"]["bench_wall_s"] = time.perf_counter() - t_start
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
    for path in shm_files:
        try:
            os.unlink(path)
        except OSError:
            pass
    if world > 1:
        drop_cache(c)
        dist.barrier()
        dist.destroy_process_group()


def probe_mode(L, r, dist, world, rank, c, lp, lanes, stream, sp, NL, B, D, train_steps, barrier, step_resident):
    """debug: in-process shard read through the real peer mappings, then sampling-only / gather-only / full loops."""
    import torch
    r.set_epoch(1)
    if os.environ.get("LGN_BENCH_PEER_DEBUG") and c.kg_bind > 1:
        peer_debug = {}
        n_dbg = min(200_000, r.capacity)
        for tag, rows in (("whole_shard", c.cap), ("first_64Ki_rows", min(c.cap, 65536))):
            dist.barrier()
            ms_dbg = r.debug_shard_read(n_dbg, rows, peers_only=True, repeats=6, stream=lp[0])
            peer_debug[tag + "_GBps"] = n_dbg * 4 * D / (ms_dbg / 1e3) / 1e9
        log("rank %d peer_debug %s" % (rank, peer_debug))
        dist.barrier()
        if os.environ["LGN_BENCH_PEER_DEBUG"] == "exit":
            return

    def ev_time(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn(n)
        for q in range(NL):
            r.wait_pipe(q, stream=sp)
        for x in lanes[1:]:
            stream.wait_stream(x)
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / n
    for nl in (1, 2, 4, 8):
        if nl > NL:
            break

        def samp(n, nl=nl):
            for i in range(n):
                r.batch_generate(L.MODE_TRAIN, B, i % train_steps, stream=lp[i % nl], pipe=i % nl)
                r.run_batch(with_features=False, stream=lp[i % nl])
        samp(8)
        log(f"rank {rank} sampling only, {nl} lanes: {ev_time(samp, 40):.4f} ms/step")
    for nl in (1, 2, 4):
        if nl > NL:
            break
        for q in range(nl):     # one sampled batch per lane, then gathers only
            r.batch_generate(L.MODE_TRAIN, B, q, stream=lp[q], pipe=q)
            r.run_batch(with_features=False, stream=lp[q])

        def gath(n, nl=nl):
            for i in range(n):
                r.select_pipe(i % nl)
                r.gather_all(stream=lp[i % nl])
        gath(4)
        log(f"rank {rank} gather only, {nl} streams: {ev_time(gath, 40):.4f} ms/batch")

    def both(n):
        for i in range(n):
            step_resident(i)
    both(8)
    both(64)
    log(f"rank {rank} full pipeline, {NL} lanes: {ev_time(both, 400):.4f} ms/step")
    if os.environ.get("LGN_NCU_RANGE"):       # ncu --replay-mode app-range: whole-range metrics under real concurrency
        def samp4(n):
            for i in range(n):
                r.batch_generate(L.MODE_TRAIN, B, i % train_steps, stream=lp[i % NL], pipe=i % NL)
                r.run_batch(with_features=False, stream=lp[i % NL])
        fn = samp4 if os.environ["LGN_NCU_RANGE"] == "samp" else both
        fn(8)
        barrier()
        torch.cuda.profiler.start()
        fn(40)
        barrier()
        torch.cuda.profiler.stop()
        return
    both(NL * 2)
    barrier()
    r.profile_enable(1024)
    both(NL * 3)
    barrier()
    if rank == 0:
        names = {0: "begin", 1: "sample", 2: "gather", 3: "end"}
        for kind, pipe, a, b in r.profile_timeline():
            log(f"lane {pipe} {names[kind]:7s} {1e3 * a:9.1f} -> {1e3 * b:9.1f} us  ({1e3 * (b - a):7.1f})")


def main():
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, the reference-style
    # server prints) write to fd 1 too, so everything but the final line is redirected to stderr
    global _real_stdout
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
