"""Host-side mirror of the reference's per-GPU pipeline objects over the C-ABI.

`Runner` plays GPURunner + GPUMemoryPool (Server.cu:167-364): it owns one lgn_ctx and
exposes the five operators (Operator.cu:10-123) under their reference roles.  The cache
planning helpers mirror GPUCache::CandidateSelection / CostModel / FillUp
(GPUCache.cu:578-826).  Every method is a thin call into liblegion_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DevArray, MappedHostArray, check, lib, _vp


class Stream:
    """a CUDA stream owned through the C-ABI (lgn_stream_create)."""

    def __init__(self, high_priority=True):
        h = C.c_void_p()
        check(lib().lgn_stream_create(C.byref(h), int(high_priority)), "lgn_stream_create")
        self.handle = h.value

    def synchronize(self):
        check(lib().lgn_stream_synchronize(C.c_void_p(self.handle)), "lgn_stream_synchronize")

    def close(self):
        if self.handle:
            lib().lgn_stream_destroy(C.c_void_p(self.handle))
            self.handle = None


def _ptr(x):
    if isinstance(x, MappedHostArray):
        return C.c_void_p(x.ptr)
    return _vp(x)


class Runner:
    def __init__(self, n_nodes, feat_dim, batch_size, fanout, device=0, part=0, rng_mode=_lib.RNG_PHILOX,
                 rng_seed=0, max_feature_rows=0, enable_hotness=False, n_lanes=0):
        cfg = _lib.Config()
        cfg.device, cfg.part, cfg.n_nodes, cfg.feat_dim = device, part, n_nodes, feat_dim
        cfg.batch_size, cfg.n_hops = batch_size, len(fanout)
        for i, f in enumerate(fanout):
            cfg.fanout[i] = f
        cfg.rng_mode, cfg.rng_seed = rng_mode, rng_seed
        cfg.max_feature_rows, cfg.enable_hotness, cfg.n_lanes = max_feature_rows, int(enable_hotness), n_lanes
        self.n_lanes = n_lanes or _lib.PIPELINE_DEPTH
        self.cfg = cfg
        self.fanout = list(fanout)
        self.handle = C.c_void_p()
        check(lib().lgn_create(C.byref(cfg), C.byref(self.handle)), "lgn_create")
        self.capacity = lib().lgn_capacity(self.handle)
        self._keep = []          # bound arrays must outlive the context
        self.pipe = 0

    # ---- storage binding (GPUNodeStorage / GPUGraphStorage) -----------------------
    def bind_seeds(self, mode, ids, labels):
        self._keep += [ids, labels]
        n = ids.shape[0]
        check(lib().lgn_bind_seeds(self.handle, mode, _ptr(ids), _ptr(labels), C.c_int32(n)), "lgn_bind_seeds")

    def bind_topology(self, indptr, indices):
        self._keep += [indptr, indices]
        check(lib().lgn_bind_topology(self.handle, _ptr(indptr), _ptr(indices)), "lgn_bind_topology")

    def bind_topology_cache(self, indptr_shards, indices_shards, slot_of, cap):
        self._keep += [indptr_shards, indices_shards, slot_of]
        n = len(indptr_shards)
        a = (C.c_void_p * max(n, 1))(*[_ptr(x).value for x in indptr_shards])
        b = (C.c_void_p * max(n, 1))(*[_ptr(x).value for x in indices_shards])
        check(lib().lgn_bind_topology_cache(self.handle, n, a, b, _ptr(slot_of), C.c_int64(cap)), "lgn_bind_topology_cache")

    def bind_features(self, features):
        self._keep.append(features)
        check(lib().lgn_bind_features(self.handle, _ptr(features)), "lgn_bind_features")

    def bind_feature_cache(self, shards, slot_of, cap):
        self._keep += [shards, slot_of]
        n = len(shards)
        a = (C.c_void_p * max(n, 1))(*[_ptr(x).value for x in shards])
        check(lib().lgn_bind_feature_cache(self.handle, n, a, _ptr(slot_of), C.c_int64(cap)), "lgn_bind_feature_cache")

    def bind_feature_cache_compact(self, shards, cmap, n_repl, cap):
        """shards addressed through a compact placement map (place_compact); cmap=None with ONE shard holding the whole
        matrix in node-id order binds it directly (no lookup)."""
        self._keep += [shards, cmap]
        n = len(shards)
        a = (C.c_void_p * max(n, 1))(*[_ptr(x).value for x in shards])
        check(lib().lgn_bind_feature_cache_compact(self.handle, n, a, _ptr(cmap) if cmap is not None else None, C.c_int64(n_repl), C.c_int64(cap)),
              "lgn_bind_feature_cache_compact")

    # ---- operators ----------------------------------------------------------------
    def batch_generate(self, mode, batch_size, counter, stream=None, pipe=None):
        """Batch_Generator::run (Operator.cu:10-32)."""
        if pipe is not None:
            self.pipe = pipe
        check(lib().lgn_batch_generate(self.handle, _vp(stream), self.pipe, mode, batch_size, counter), "lgn_batch_generate")

    def batch_from_host(self, seeds, labels=None, step=0, stream=None, pipe=None):
        if pipe is not None:
            self.pipe = pipe
        seeds = np.ascontiguousarray(seeds, np.int32) if isinstance(seeds, np.ndarray) else seeds
        n = seeds.shape[0]
        check(lib().lgn_batch_from_host(self.handle, _vp(stream), self.pipe, _ptr(seeds), _ptr(labels), C.c_int32(n),
                                        C.c_uint32(step)), "lgn_batch_from_host")

    def sample_hop(self, hop, is_presc=False, stream=None):
        """Random_Sampler::run (Operator.cu:34-56)."""
        check(lib().lgn_sample_hop(self.handle, _vp(stream), hop, int(is_presc)), "lgn_sample_hop")

    def gather_segment(self, segment, stream=None):
        """Feature_Extractor::run (Operator.cu:58-78)."""
        check(lib().lgn_gather_segment(self.handle, _vp(stream), segment), "lgn_gather_segment")

    def gather_segments(self, first, count, stream=None):
        check(lib().lgn_gather_segments(self.handle, _vp(stream), first, count), "lgn_gather_segments")

    def gather_all(self, stream=None):
        """every feature-extraction launch of the current slot's batch, as run_batch issues them."""
        check(lib().lgn_gather_batch(self.handle, _vp(stream)), "lgn_gather_batch")

    def launches_per_batch(self, with_features=True):
        return int(lib().lgn_launches_per_batch(self.handle, int(with_features)))

    def gather_kernel_name(self):
        return lib().lgn_gather_kernel_name(self.handle).decode()

    def finish_batch(self, is_presc=False, stream=None):
        """Cache_Planner::run + Cache_Updater::run (Operator.cu:80-123)."""
        check(lib().lgn_finish_batch(self.handle, _vp(stream), int(is_presc)), "lgn_finish_batch")

    def run_batch(self, with_features=True, is_presc=False, stream=None):
        """GPURunner::RunOnce / RunPreSc minus the IPC handshake (Server.cu:284-328)."""
        check(lib().lgn_run_batch(self.handle, _vp(stream), int(with_features), int(is_presc)), "lgn_run_batch")

    def set_epoch(self, epoch, step_offset=0):
        """position of the following batches in the Philox stream (counter words 1 and 3)."""
        check(lib().lgn_set_epoch(self.handle, C.c_uint32(epoch), C.c_uint32(step_offset)), "lgn_set_epoch")

    def set_part(self, part):
        check(lib().lgn_set_part(self.handle, C.c_int32(part)), "lgn_set_part")

    def set_dedup_capacity(self, expected_unique):
        check(lib().lgn_set_dedup_capacity(self.handle, C.c_int64(expected_unique)), "lgn_set_dedup_capacity")

    def select_pipe(self, pipe):
        self.pipe = pipe
        check(lib().lgn_select_pipe(self.handle, pipe), "lgn_select_pipe")

    def wait_pipe(self, pipe, stream=None):
        check(lib().lgn_wait_pipe(self.handle, _vp(stream), pipe), "lgn_wait_pipe")

    def sync_pipe(self, pipe):
        check(lib().lgn_sync_pipe(self.handle, pipe), "lgn_sync_pipe")

    # ---- results ------------------------------------------------------------------
    def view(self, pipe=None):
        v = _lib.BatchView()
        check(lib().lgn_batch_buffers(self.handle, self.pipe if pipe is None else pipe, C.byref(v)), "lgn_batch_buffers")
        return v

    def read_counters(self, stream=None, pipe=None):
        nc = np.zeros(16, np.int32)
        ec = np.zeros(16, np.int32)
        check(lib().lgn_read_counters(self.handle, _vp(stream), self.pipe if pipe is None else pipe, _vp(nc), _vp(ec)),
              "lgn_read_counters")
        return nc, ec

    def fetch(self, with_features=True, stream=None):
        """copy the whole batch to host numpy arrays (tests only)."""
        nc, ec = self.read_counters(stream)
        v = self.view()
        n_hops = self.cfg.n_hops
        total = int(nc[0])
        n_e = int(ec[0])
        cap = int(v.capacity)
        get = lambda p, n, dt=np.int32: DevArray((cap,), dt, ptr=p, owner=False).numpy(n)
        out = dict(nc=nc, ec=ec, sampled_ids=get(v.ids, total), agg_src_ids=get(v.agg_src_ids, n_e),
                   agg_dst_ids=get(v.agg_dst_ids, n_e), agg_src_off=get(v.agg_src, n_e), agg_dst_off=get(v.agg_dst, n_e),
                   labels=DevArray((self.cfg.batch_size,), np.int32, ptr=v.labels, owner=False).numpy(int(nc[4])))
        if with_features and v.features:
            rows = min(total, int(v.max_rows))
            out["features"] = DevArray((int(v.max_rows), self.cfg.feat_dim), np.float32, ptr=v.features, owner=False).numpy(rows)
        out["n_hops"] = n_hops
        return out

    def tier_counts(self, reset=False, stream=None):
        out = (C.c_int64 * 3)()
        check(lib().lgn_tier_counts(self.handle, _vp(stream), out, int(reset)), "lgn_tier_counts")
        return list(out)

    def status(self, stream=None):
        return lib().lgn_status(self.handle, _vp(stream))

    def debug_shard_read(self, n_rows, rows_per_shard, peers_only=True, repeats=6, pipe=0, stream=None):
        """diagnostic: average ms of a plain random row read out of the bound cache shards (lgn_debug_shard_read)."""
        ms = C.c_double()
        check(lib().lgn_debug_shard_read(self.handle, _vp(stream), C.c_int32(pipe), C.c_int64(n_rows), C.c_int64(rows_per_shard),
                                         C.c_int32(1 if peers_only else 0), C.c_int32(repeats), C.byref(ms)), "lgn_debug_shard_read")
        return ms.value

    def hotness(self):
        a, b = C.c_void_p(), C.c_void_p()
        check(lib().lgn_hotness(self.handle, C.byref(a), C.byref(b)), "lgn_hotness")
        n = self.cfg.n_nodes
        return DevArray((n,), np.uint32, ptr=a.value, owner=False), DevArray((n,), np.uint32, ptr=b.value, owner=False)

    def max_ids(self, stream=None):
        return lib().lgn_max_ids(self.handle, _vp(stream))

    def profile_enable(self, max_records):
        check(lib().lgn_profile_enable(self.handle, C.c_int32(max_records)), "lgn_profile_enable")

    def profile_collect(self):
        ms = (C.c_double * 4)()
        calls = (C.c_int64 * 4)()
        check(lib().lgn_profile_collect(self.handle, ms, calls), "lgn_profile_collect")
        return list(ms), list(calls)

    def profile_timeline(self, max_records=4096):
        rows = (C.c_double * (4 * max_records))()
        n = C.c_int32()
        check(lib().lgn_profile_timeline(self.handle, rows, max_records, C.byref(n)), "lgn_profile_timeline")
        return [(int(rows[4 * i]), int(rows[4 * i + 1]), rows[4 * i + 2], rows[4 * i + 3]) for i in range(n.value)]

    def close(self):
        if self.handle:
            lib().lgn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- planner (GPUCache::CandidateSelection / CostModel / FillUp) ------------------------
def hot_order(counts, want_sorted=False, stream=None):
    n = counts.shape[0]
    order = DevArray((n,), np.int32)
    sorted_counts = DevArray((n,), np.uint32) if want_sorted else None
    check(lib().lgn_hot_order(_ptr(counts), C.c_int64(n), _ptr(order), _ptr(sorted_counts), _vp(stream)), "lgn_hot_order")
    return (order, sorted_counts) if want_sorted else order


def place(order, cap, kg, stream=None):
    n = order.shape[0]
    slot_of = DevArray((n,), np.int32)
    check(lib().lgn_place(_ptr(order), C.c_int64(n), C.c_int64(cap), C.c_int32(kg), _ptr(slot_of), _vp(stream)), "lgn_place")
    return slot_of


def fill_feature_shard(order, cap, kg, j, features, dim, stream=None, out=None):
    n = order.shape[0]
    shard = out if out is not None else DevArray.zeros((cap, dim), np.float32)
    check(lib().lgn_fill_feature_shard(_ptr(order), C.c_int64(n), C.c_int64(cap), C.c_int32(kg), C.c_int32(j),
                                       _ptr(features), C.c_int32(dim), _ptr(shard), _vp(stream)), "lgn_fill_feature_shard")
    return shard


def shared_alloc(shape, dtype):
    """device array other processes can map (lgn_shared_alloc): returns (array, fd, mapped_bytes); close fd after the exchange,
    release the array with shared_free."""
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    p, fd, mapped = C.c_void_p(), C.c_int32(-1), C.c_int64(0)
    check(lib().lgn_shared_alloc(C.byref(p), C.c_int64(max(1, nbytes)), C.byref(fd), C.byref(mapped)), "lgn_shared_alloc")
    return DevArray(shape, dtype, ptr=p.value, owner=False), fd.value, mapped.value


def shared_import(fd, mapped_bytes, shape, dtype):
    """map another process's shared_alloc array for THIS process's current device (lgn_shared_import)."""
    p = C.c_void_p()
    check(lib().lgn_shared_import(C.c_int32(fd), C.c_int64(mapped_bytes), C.byref(p)), "lgn_shared_import")
    return DevArray(tuple(int(x) for x in np.atleast_1d(shape)), dtype, ptr=p.value, owner=False)


def shared_free(arr):
    check(lib().lgn_shared_free(C.c_void_p(arr.ptr)), "lgn_shared_free")


def place_hybrid(order, cap, kg, n_repl, my_part, stream=None):
    """B200 extension: n_repl hottest ranks replicated on every GPU, the rest partitioned (lgn_place_hybrid)."""
    n = order.shape[0]
    slot_of = DevArray((n,), np.int32)
    check(lib().lgn_place_hybrid(_ptr(order), C.c_int64(n), C.c_int64(cap), C.c_int32(kg), C.c_int64(n_repl), C.c_int32(my_part),
                                 _ptr(slot_of), _vp(stream)), "lgn_place_hybrid")
    return slot_of


def fill_feature_shard_hybrid(order, cap, kg, j, n_repl, features, dim, stream=None, out=None):
    n = order.shape[0]
    shard = out if out is not None else DevArray.zeros((cap, dim), np.float32)
    check(lib().lgn_fill_feature_shard_hybrid(_ptr(order), C.c_int64(n), C.c_int64(cap), C.c_int32(kg), C.c_int32(j), C.c_int64(n_repl),
                                              _ptr(features), C.c_int32(dim), _ptr(shard), _vp(stream)), "lgn_fill_feature_shard_hybrid")
    return shard


def place_compact(order, n_repl, n_part, stream=None):
    """B200 extension: compact, L2-resident placement map (lgn_place_compact): the n_repl hottest ranks replicated, the next
    n_part partitioned, rows of a class in node-id order.  -> uint32[records * 8]"""
    n = order.shape[0]
    cmap = DevArray((int(lib().lgn_cmap_bytes(C.c_int64(n))) // 4,), np.uint32)
    check(lib().lgn_place_compact(_ptr(order), C.c_int64(n), C.c_int64(n_repl), C.c_int64(n_part), _ptr(cmap), _vp(stream)), "lgn_place_compact")
    return cmap


def fill_feature_shard_compact(cmap, n, n_repl, kg, j, features, dim, cap, stream=None, out=None):
    shard = out if out is not None else DevArray.zeros((cap, dim), np.float32)
    check(lib().lgn_fill_feature_shard_compact(_ptr(cmap), C.c_int64(n), C.c_int64(n_repl), C.c_int32(kg), C.c_int32(j), _ptr(features),
                                               C.c_int32(dim), _ptr(shard), C.c_int64(cap), _vp(stream)), "lgn_fill_feature_shard_compact")
    return shard


def fill_topo_shard(order, cap, kg, j, indptr, indices, stream=None):
    n = order.shape[0]
    ip = DevArray((cap + 1,), np.int64)
    cnt = C.c_int64()
    args = (_ptr(order), C.c_int64(n), C.c_int64(cap), C.c_int32(kg), C.c_int32(j), _ptr(indptr), _ptr(indices), _ptr(ip))
    check(lib().lgn_fill_topo_shard(*args, None, C.byref(cnt), _vp(stream)), "lgn_fill_topo_shard(size)")
    ix = DevArray((max(cnt.value, 1),), np.int32)
    check(lib().lgn_fill_topo_shard(*args, _ptr(ix), C.byref(cnt), _vp(stream)), "lgn_fill_topo_shard(fill)")
    return ip, ix, cnt.value


def cost_model(af_sorted, at_sorted, qt, indptr, dim, cache_memory, kg, topo_trans, max_ids, train_step):
    n = qt.shape[0]
    mi = (C.c_int32 * len(max_ids))(*max_ids)
    ncap, ecap = C.c_int32(), C.c_int32()
    check(lib().lgn_cost_model(_ptr(af_sorted), _ptr(at_sorted), _ptr(qt), _ptr(indptr), C.c_int64(n), C.c_int32(dim),
                               C.c_int64(cache_memory), C.c_int32(kg), C.c_uint64(topo_trans), mi, C.c_int32(train_step),
                               C.byref(ncap), C.byref(ecap), None), "lgn_cost_model")
    return ncap.value, ecap.value


def plan_hybrid(af_sorted, dim, budget_bytes, kg, bw_local, bw_peer, bw_host, prior=0.5, stream=None):
    """B200 placement model (lgn_plan_hybrid): -> (n_repl, cap, estimated cost)."""
    n = af_sorted.shape[0]
    n_repl, cap, cost = C.c_int64(), C.c_int64(), C.c_double()
    check(lib().lgn_plan_hybrid(_ptr(af_sorted), C.c_int64(n), C.c_int32(dim), C.c_int64(int(budget_bytes)), C.c_int32(kg),
                                C.c_double(bw_local), C.c_double(bw_peer), C.c_double(bw_host), C.c_double(prior), C.byref(n_repl), C.byref(cap),
                                C.byref(cost), _vp(stream)), "lgn_plan_hybrid")
    return n_repl.value, cap.value, cost.value


def coordinate(n_train, n_valid, n_test, batch, epochs):
    s = _lib.Steps()
    P = len(n_train)
    arr = lambda v: (C.c_int32 * P)(*v)
    check(lib().lgn_coordinate(arr(n_train), arr(n_valid), arr(n_test), C.c_int32(P), C.c_int32(batch), C.c_int32(epochs),
                               C.byref(s)), "lgn_coordinate")
    return s
