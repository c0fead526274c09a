"""ctypes/numpy front end of the CPU oracle (oracle/legion_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(legion-1_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblegion_oracle.so")

RNG_MINSTD, RNG_PHILOX = 0, 1


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("legion_oracle.c", "legion_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.lgo_minstd_pow.restype = C.c_uint32
        _lib.lgo_minstd_pow.argtypes = [C.c_uint64]
        _lib.lgo_minstd_pick.restype = C.c_int32
        _lib.lgo_minstd_pick.argtypes = [C.c_uint64, C.c_int32]
        _lib.lgo_philox_pick.restype = C.c_int32
        _lib.lgo_philox_pick.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32]
        _lib.lgo_batch_generate.restype = C.c_int32
        _lib.lgo_fill_topo_shard.restype = C.c_int64
        _lib.lgo_mode_of_step.restype = C.c_int32
        _lib.lgo_local_batch_id.restype = C.c_int32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SampleArgs(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("indptr", C.c_void_p), ("indices", C.c_void_p),
        ("n_hops", C.c_int32), ("fanout", C.c_void_p), ("rng_mode", C.c_int32),
        ("rng_seed", C.c_uint64), ("step", C.c_uint32), ("epoch", C.c_uint32),
        ("capacity", C.c_int64), ("n_seeds", C.c_int32),
        ("sampled_ids", C.c_void_p), ("agg_src_ids", C.c_void_p), ("agg_dst_ids", C.c_void_p),
        ("agg_src_off", C.c_void_p), ("agg_dst_off", C.c_void_p),
        ("nc", C.c_void_p), ("ec", C.c_void_p), ("position_map", C.c_void_p),
        ("topo_hotness", C.c_void_p), ("node_hotness", C.c_void_p), ("n_threads", C.c_int32),
    ]


class Steps(C.Structure):
    _fields_ = [("train_step", C.c_int32), ("valid_step", C.c_int32), ("test_step", C.c_int32),
                ("max_step", C.c_int32), ("valid_batch", C.c_int32 * 8), ("test_batch", C.c_int32 * 8)]


def minstd_pick(idx, deg):
    return lib().lgo_minstd_pick(int(idx), int(deg))


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().lgo_philox4x32_10(c, k, o)
    return list(o)


def philox_pick(idx, hop, step, seed, deg, epoch=0):
    return lib().lgo_philox_pick(C.c_uint64(int(idx)), C.c_uint32(int(epoch)), C.c_uint32(int(hop)), C.c_uint32(int(step)),
                                 C.c_uint64(int(seed)), C.c_int32(int(deg)))


def capacity_for(batch, fanout):
    """Server.cu:184-196: B*(1 + f1 + f1*f2 + ...)."""
    tot, cur = batch, batch
    for f in fanout:
        cur *= f
        tot += cur
    return tot


def batch_generate(all_ids, all_labels, batch_size, counter):
    ids = np.full(batch_size, -1, np.int32)
    labels = np.full(batch_size, -1, np.int32)
    n = lib().lgo_batch_generate(_p(all_ids), _p(all_labels), C.c_int32(len(all_ids)),
                                 C.c_int32(batch_size), C.c_int32(counter), _p(ids), _p(labels))
    return ids[:n].copy(), labels[:n].copy()


class Sampler:
    """k-hop sampler + dedup + relabel (Kernels.cu:342-463)."""

    def __init__(self, indptr, indices, fanout, rng_mode=RNG_PHILOX, rng_seed=0, capacity=None, n_threads=1):
        self.indptr = np.ascontiguousarray(indptr, np.int64)
        self.indices = np.ascontiguousarray(indices, np.int32)
        self.n = len(self.indptr) - 1
        self.fanout = np.asarray(fanout, np.int32)
        self.rng_mode, self.rng_seed, self.n_threads = rng_mode, rng_seed, n_threads
        self.pos = np.full(self.n, -1, np.int32)
        self.capacity = capacity
        self.topo_hotness = None
        self.node_hotness = None

    def enable_hotness(self):
        self.topo_hotness = np.zeros(self.n, np.uint32)
        self.node_hotness = np.zeros(self.n, np.uint32)

    def sample(self, seeds, step=0, epoch=0):
        seeds = np.asarray(seeds, np.int32)
        B = len(seeds)
        cap = self.capacity or capacity_for(max(B, 1), self.fanout.tolist())
        out = {k: np.full(cap, -1, np.int32) for k in
               ("sampled_ids", "agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off")}
        out["sampled_ids"][:B] = seeds
        nc = np.zeros(16, np.int32)
        ec = np.zeros(16, np.int32)
        a = SampleArgs(self.n, _p(self.indptr), _p(self.indices), len(self.fanout), _p(self.fanout),
                       self.rng_mode, self.rng_seed, step, epoch, cap, B,
                       _p(out["sampled_ids"]), _p(out["agg_src_ids"]), _p(out["agg_dst_ids"]),
                       _p(out["agg_src_off"]), _p(out["agg_dst_off"]), _p(nc), _p(ec), _p(self.pos),
                       _p(self.topo_hotness), _p(self.node_hotness), self.n_threads)
        rc = lib().lgo_sample_batch(C.byref(a))
        if rc != 0:
            raise RuntimeError(f"lgo_sample_batch failed rc={rc}")
        out["nc"], out["ec"] = nc, ec
        return out


def draw_hop(indptr, indices, frontier, f, rng_mode=RNG_MINSTD, rng_seed=0, hop=0, step=0, epoch=0):
    frontier = np.ascontiguousarray(frontier, np.int32)
    out = np.empty(len(frontier) * f, np.int32)
    lib().lgo_draw_hop(_p(indptr), _p(indices), _p(frontier), C.c_int64(len(frontier)), C.c_int32(f), C.c_int32(rng_mode),
                       C.c_uint64(rng_seed), C.c_uint32(hop), C.c_uint32(step), C.c_uint32(epoch), _p(out))
    return out


def hot_order(counts):
    counts = np.ascontiguousarray(counts, np.uint32)
    order = np.empty(len(counts), np.int32)
    lib().lgo_hot_order(_p(counts), C.c_int64(len(counts)), _p(order))
    return order


def place(order, cap, kg):
    slot = np.empty(len(order), np.int32)
    lib().lgo_place(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), _p(slot))
    return slot


def fill_feature_shard(order, cap, kg, j, features):
    features = np.ascontiguousarray(features, np.float32)
    shard = np.zeros((cap, features.shape[1]), np.float32)
    lib().lgo_fill_feature_shard(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), C.c_int32(j),
                                 _p(features), C.c_int32(features.shape[1]), _p(shard))
    return shard


def place_hybrid(order, cap, kg, n_repl, my_part):
    slot = np.empty(len(order), np.int32)
    lib().lgo_place_hybrid(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), C.c_int64(n_repl), C.c_int32(my_part), _p(slot))
    return slot


def fill_feature_shard_hybrid(order, cap, kg, j, n_repl, features):
    features = np.ascontiguousarray(features, np.float32)
    shard = np.zeros((cap, features.shape[1]), np.float32)
    lib().lgo_fill_feature_shard_hybrid(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), C.c_int32(j), C.c_int64(n_repl),
                                        _p(features), C.c_int32(features.shape[1]), _p(shard))
    return shard


CMAP_NODES = 96


def place_compact(order, n_repl, n_part, kg, my_part, cap):
    """compact placement (include/legion_b200.h: lgn_place_compact), restated with numpy: classes by hotness rank
    (GPUCache.cu:103-108 decides WHICH rows are cached), rows of a class in node-id order.
    -> (record words uint32[records * 8], the equivalent int32 slot table part*cap+row / -1 of GPU my_part)"""
    n = len(order)
    is_r = np.zeros(n, bool)
    is_p = np.zeros(n, bool)
    is_r[order[:n_repl]] = True
    is_p[order[n_repl:n_repl + n_part]] = True
    rank_r = np.cumsum(is_r) - 1
    q = np.cumsum(is_p) - 1
    slot = np.full(n, -1, np.int64)
    slot[is_r] = my_part * cap + rank_r[is_r]
    slot[is_p] = (q[is_p] % kg) * cap + n_repl + q[is_p] // kg
    n_rec = (n + CMAP_NODES - 1) // CMAP_NODES
    pad = n_rec * CMAP_NODES - n
    words = np.zeros((n_rec, 8), np.uint32)
    for plane, col in ((is_r, 2), (is_p, 5)):
        bits = np.concatenate([plane, np.zeros(pad, bool)]).reshape(n_rec, 3, 32)
        words[:, col:col + 3] = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32)
        per_rec = bits.reshape(n_rec, 96).sum(axis=1)
        words[:, 0 if col == 2 else 1] = (np.cumsum(per_rec) - per_rec).astype(np.uint32)
    return words.reshape(-1), slot.astype(np.int32)


def fill_feature_shard_compact(slot_of, cap, j, features):
    """shard j of a placement given as a slot table (rows another GPU's table marks local for a replicated node are the
    same rows on every GPU)."""
    features = np.ascontiguousarray(features, np.float32)
    shard = np.zeros((cap, features.shape[1]), np.float32)
    mine = (slot_of >= 0) & (slot_of // cap == j)
    shard[slot_of[mine] % cap] = features[mine]
    return shard


def fill_topo_shard(order, cap, kg, j, indptr, indices):
    ip = np.zeros(cap + 1, np.int64)
    n = lib().lgo_fill_topo_shard(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), C.c_int32(j),
                                  _p(indptr), _p(indices), _p(ip), None)
    ix = np.zeros(max(n, 1), np.int32)
    lib().lgo_fill_topo_shard(_p(order), C.c_int64(len(order)), C.c_int64(cap), C.c_int32(kg), C.c_int32(j),
                              _p(indptr), _p(indices), _p(ip), _p(ix))
    return ip, ix[:n]


def cost_model(af, at, qt, indptr, dim, cache_memory, kg, topo_trans, max_ids, train_step):
    af = np.ascontiguousarray(af, np.uint64)
    at = np.ascontiguousarray(at, np.uint64)
    qt = np.ascontiguousarray(qt, np.int32)
    mi = np.ascontiguousarray(max_ids, np.int32)
    ncap, ecap, best = C.c_int32(), C.c_int32(), C.c_int32()
    lib().lgo_cost_model(_p(af), _p(at), _p(qt), _p(indptr), C.c_int64(len(qt)), C.c_int32(dim),
                         C.c_int64(cache_memory), C.c_int32(kg), C.c_uint64(topo_trans), _p(mi),
                         C.c_int32(train_step), C.byref(ncap), C.byref(ecap), C.byref(best))
    return ncap.value, ecap.value, best.value


def gather(sampled_ids, off, cnt, slot_of, cap, shards, host_features, out, n_threads=1, tiers=False):
    host_features = np.ascontiguousarray(host_features, np.float32)
    n, dim = host_features.shape
    ns = len(shards) if shards else 0
    ptrs = (C.c_void_p * max(ns, 1))(*[s.ctypes.data for s in (shards or [])])
    tr = np.zeros(ns + 1, np.int64) if tiers else None
    lib().lgo_gather(_p(sampled_ids), C.c_int32(off), C.c_int32(cnt), _p(slot_of), C.c_int64(max(cap, 1)), ptrs,
                     _p(host_features), C.c_int64(n), C.c_int32(dim), _p(out), _p(tr), C.c_int32(ns),
                     C.c_int32(n_threads))
    return tr


def coordinate(n_train, n_valid, n_test, batch, epochs):
    s = Steps()
    P = len(n_train)
    arr = lambda v: (C.c_int32 * P)(*v)
    lib().lgo_coordinate(arr(n_train), arr(n_valid), arr(n_test), C.c_int32(P), C.c_int32(batch), C.c_int32(epochs),
                         C.byref(s))
    return s


def mode_of_step(s, epochs, g):
    return lib().lgo_mode_of_step(C.byref(s), C.c_int32(epochs), C.c_int32(g))


def local_batch_id(s, epochs, g):
    return lib().lgo_local_batch_id(C.byref(s), C.c_int32(epochs), C.c_int32(g))
