"""A/B against the reference itself: its own kernels (Kernels.cu, GPUCache.cu, ...) compiled
unmodified for sm_100a by oracle/ref_harness and driven below its Server class.  This is what
pins the oracle: the reference's output order inside a hop is atomic-arrival order
(Kernels.cu:418-445), so buffers are compared canonicalised -- sorted id sets
per hop segment, sorted raw-id edge multisets per hop, features keyed by node id -- while
counters, hotness histograms, hot orders, capacities and cache contents compare exactly."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_legion.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_legion.so not built (needs /root/reference at build time)")
    lib = C.CDLL(REF_SO)
    lib.ref_create.restype = C.c_void_p
    lib.ref_capacity.restype = C.c_int64
    return lib


def _canon_edges(src, dst):
    k = np.stack([dst.astype(np.int64), src.astype(np.int64)], 1)
    return k[np.lexsort((k[:, 1], k[:, 0]))]


def _check_batch_against_replay(O, d, B, f1, f2, ids, s_ids, d_ids, nc, ec, who="reference"):
    """One batch of the reference, re-derived from the oracle's draw function.  Hop 1 is a pure function
    of the seeds; hop 2's slot index -- and with it the minstd draw -- depends on the ORDER of the hop-1
    edge list, which the reference fills in atomic-arrival order (Kernels.cu:371-373, 418-445), so hop 2
    is replayed over the reference's own hop-1 order."""
    total, e1, e2 = int(nc[9]), int(ec[3]), int(ec[4])
    seeds = ids[:B]
    h1 = O.draw_hop(d.indptr, d.indices, seeds, f1, O.RNG_MINSTD, hop=0)
    v1 = h1 >= 0
    assert e1 == int(v1.sum()), who
    assert np.array_equal(_canon_edges(s_ids[:e1], d_ids[:e1]), _canon_edges(h1[v1], np.repeat(seeds, f1)[v1])), who
    frontier = s_ids[:e1]
    h2 = O.draw_hop(d.indptr, d.indices, frontier, f2, O.RNG_MINSTD, hop=1)
    v2 = h2 >= 0
    assert e2 - e1 == int(v2.sum()), who
    assert np.array_equal(_canon_edges(s_ids[e1:e2], d_ids[e1:e2]), _canon_edges(h2[v2], np.repeat(frontier, f2)[v2])), who
    seg1 = np.setdiff1d(np.unique(h1[v1]), seeds)
    seg2 = np.setdiff1d(np.setdiff1d(np.unique(h2[v2]), seeds), seg1)
    assert list(nc[:10]) == [total, 0, e2 - e1, 0, B, B, len(seg1), B + len(seg1), len(seg2), B + len(seg1) + len(seg2)], (who, nc)
    assert list(ec[:5]) == [e2, 0, e1, e1, e2], (who, ec)
    assert np.array_equal(np.sort(ids[B:B + len(seg1)]), seg1) and np.array_equal(np.sort(ids[B + len(seg1):total]), seg2), who


def test_reference_kernels_agree_with_oracle_and_cuda_path(ref):
    import legion_b200 as L
    from oracle import oracle as O
    d = L.synth.make_dataset(20_000, 12.0, 32, n_class=7)
    B, f1, f2 = 256, 25, 10
    train = d.train_ids[:2000].astype(np.int32)
    labels = d.labels[train]
    cache_mem = 1_500_000        # restricted so that the cost model has to split topology vs features
    h = C.c_void_p(ref.ref_create(_p(d.indptr), _p(d.indices), C.c_int32(d.n_nodes), C.c_int64(d.n_edges), _p(d.features),
                                  C.c_int32(d.dim), _p(train), _p(labels), C.c_int32(len(train)), C.c_int32(B),
                                  C.c_int32(f1), C.c_int32(f2), C.c_int64(cache_mem)))
    train_step = (len(train) - 1) // B
    cap = ref.ref_capacity(h)
    smp = O.Sampler(d.indptr, d.indices, [f1, f2], rng_mode=O.RNG_MINSTD)
    run = L.Runner(d.n_nodes, d.dim, B, [f1, f2], rng_mode=L.RNG_MINSTD, enable_hotness=True)
    ipd, ixd = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices)
    run.bind_topology(ipd, ixd)
    run.bind_features(L.DevArray.from_numpy(d.features))
    run.bind_seeds(L.MODE_TRAIN, L.DevArray.from_numpy(train), L.DevArray.from_numpy(labels))

    # ---- presampling epoch (Kernels.cu:468-564, GPUCache.cu:227-235, 294-296): every batch re-derived from the
    # oracle's draws; the reference's hotness histograms must be exactly the sums over those batches
    e_node, e_topo, e_max = np.zeros(d.n_nodes, np.uint64), np.zeros(d.n_nodes, np.uint64), 0
    for it in range(train_step):
        ids = np.zeros(cap, np.int32)
        s_ids, d_ids = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        nc, ec = np.zeros(16, np.int32), np.zeros(16, np.int32)
        ref.ref_presample_batch(h, C.c_int32(it), _p(ids), _p(s_ids), _p(d_ids), _p(nc), _p(ec))
        assert np.array_equal(ids[:B], train[it * B:(it + 1) * B])
        _check_batch_against_replay(O, d, B, f1, f2, ids, s_ids, d_ids, nc, ec)
        np.add.at(e_node, ids[:nc[9]], 1)
        np.add.at(e_topo, d_ids[:ec[4]], 1)
        e_max = max(e_max, int(nc[9]))
        # the oracle's canonical batch and the CUDA path satisfy the same replay relations
        o = smp.sample(train[it * B:(it + 1) * B], step=it)
        _check_batch_against_replay(O, d, B, f1, f2, o["sampled_ids"], o["agg_src_ids"], o["agg_dst_ids"], o["nc"], o["ec"], "oracle")
        run.batch_generate(L.MODE_TRAIN, B, it, pipe=it % 2)
        run.run_batch(with_features=False, is_presc=True)
        g = run.fetch(with_features=False)
        _check_batch_against_replay(O, d, B, f1, f2, g["sampled_ids"], g["agg_src_ids"], g["agg_dst_ids"], g["nc"], g["ec"], "cuda")
        # hop 1 does not depend on arrival order: identical across all three
        e1 = int(ec[3])
        assert e1 == o["ec"][3] == g["ec"][3] and nc[6] == o["nc"][6] == g["nc"][6]
        assert np.array_equal(np.sort(ids[B:nc[7]]), np.sort(o["sampled_ids"][B:nc[7]]))
    r_node, r_topo = np.zeros(d.n_nodes, np.uint64), np.zeros(d.n_nodes, np.uint64)
    r_max = C.c_int32()
    ref.ref_hotness(h, _p(r_node), _p(r_topo), C.byref(r_max))
    assert np.array_equal(r_node, e_node) and np.array_equal(r_topo, e_topo) and r_max.value == e_max

    # ---- candidate selection, cost model, fill-up (GPUCache.cu:578-826) on the reference's own histograms
    qf, qt = np.zeros(d.n_nodes, np.int32), np.zeros(d.n_nodes, np.int32)
    ncap, ecap = C.c_int32(), C.c_int32()
    shard = np.zeros((d.n_nodes, d.dim), np.float32)
    topo_trans = 1_000_000
    ref.ref_plan(h, C.c_uint64(topo_trans), _p(qf), _p(qt), C.byref(ncap), C.byref(ecap), _p(shard), C.c_int64(d.n_nodes))
    n32, t32 = r_node.astype(np.uint32), r_topo.astype(np.uint32)
    o_qf, o_qt = O.hot_order(n32), O.hot_order(t32)
    assert np.array_equal(qf, o_qf), "feature hot order (count desc, id asc) differs from thrust::sort_by_key(greater)"
    assert np.array_equal(qt, o_qt)
    o_caps = O.cost_model(n32[o_qf], t32[o_qt], o_qt, d.indptr, d.dim, cache_mem, 1, topo_trans, [e_max], train_step)
    assert (ncap.value, ecap.value) == o_caps[:2], ((ncap.value, ecap.value), o_caps)
    assert 1 < ncap.value < d.n_nodes and 1 < ecap.value < d.n_nodes   # the restricted budget really splits
    o_shard = O.fill_feature_shard(o_qf, ncap.value, 1, 0, d.features)
    assert np.array_equal(shard[:ncap.value].view(np.uint32), o_shard.view(np.uint32))      # cache contents bit-exact
    # new path: planner on device, fed with the same histograms
    nh, th = L.DevArray.from_numpy(n32), L.DevArray.from_numpy(t32)
    d_qf, d_af = L.hot_order(nh, want_sorted=True)
    d_qt, d_at = L.hot_order(th, want_sorted=True)
    assert np.array_equal(d_qf.numpy(), qf) and np.array_equal(d_qt.numpy(), qt)
    assert L.cost_model(d_af, d_at, d_qt, ipd, d.dim, cache_mem, 1, topo_trans, [e_max], train_step) == (ncap.value, ecap.value)
    d_shard = L.fill_feature_shard(d_qf, ncap.value, 1, 0, L.DevArray.from_numpy(d.features), d.dim)
    assert np.array_equal(d_shard.numpy().view(np.uint32), shard[:ncap.value].view(np.uint32))
    slot = L.place(d_qf, ncap.value, 1)
    run.bind_feature_cache([d_shard], slot, ncap.value)
    tslot = L.place(d_qt, ecap.value, 1)
    tip, tix, _ = L.fill_topo_shard(d_qt, ecap.value, 1, 0, ipd, ixd)
    run.bind_topology_cache([tip], [tix], tslot, ecap.value)

    # ---- steady-state batches over the filled caches (Server.cu:301-328): cuckoo lookups, cache hits and misses
    for it in range(3):
        ids = np.zeros(cap, np.int32); lab = np.zeros(B, np.int32)
        s_ids, d_ids, s_off, d_off = (np.zeros(cap, np.int32) for _ in range(4))
        nc, ec = np.zeros(16, np.int32), np.zeros(16, np.int32)
        feats = np.zeros((cap, d.dim), np.float32)
        ref.ref_train_batch(h, C.c_int32(it), _p(ids), _p(lab), _p(s_ids), _p(d_ids), _p(s_off), _p(d_off), _p(nc), _p(ec), _p(feats))
        _check_batch_against_replay(O, d, B, f1, f2, ids, s_ids, d_ids, nc, ec)
        total, e2 = int(nc[9]), int(ec[4])
        # the reference's own relabelling is consistent with its id order, features follow the ids
        assert np.array_equal(ids[s_off[:e2]], s_ids[:e2]) and np.array_equal(ids[d_off[:e2]], d_ids[:e2])
        assert np.array_equal(feats[:total].view(np.uint32), d.features[ids[:total]].view(np.uint32))
        assert np.array_equal(lab[:nc[4]], labels[it * B:(it + 1) * B])
        # CUDA path over its own caches: same relations, features bit-exact per node id
        run.batch_generate(L.MODE_TRAIN, B, it, pipe=it % 2)
        run.run_batch(with_features=True)
        g = run.fetch()
        _check_batch_against_replay(O, d, B, f1, f2, g["sampled_ids"], g["agg_src_ids"], g["agg_dst_ids"], g["nc"], g["ec"], "cuda")
        assert np.array_equal(g["features"].view(np.uint32), d.features[g["sampled_ids"]].view(np.uint32))
        assert np.array_equal(g["labels"], lab[:nc[4]]) and np.array_equal(g["sampled_ids"][:B], ids[:B])
        assert g["ec"][3] == ec[3] and g["nc"][6] == nc[6]
    tiers = run.tier_counts()
    assert tiers[0] > 0 and tiers[2] > 0        # both cache hits and host misses were exercised
    run.close()
