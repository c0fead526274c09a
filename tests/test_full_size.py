"""Full BASELINE sizes (configs[1]: ogbn-products-shaped, 2,449,029 nodes, ~61.9 M edges, 100-d, batch 8000,
fanout [25,10]) on the GPU: the oracle is fast enough to compare whole batches bit-exactly, plus the
size-independent properties of the path (relabelling round trip, uniqueness, counter algebra, idempotence,
features against their closed form, hotness sums)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import legion_b200 as L
    cfg = L.synth.CONFIGS["C2"]
    return L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"]), cfg


def _closed_form_features(ids, dim):
    i = ids.astype(np.uint64)[:, None]
    j = np.arange(dim, dtype=np.uint64)[None, :]
    v = (i * np.uint64(2654435761) + j * np.uint64(40503)) & np.uint64(0x7FFFFF)
    return (v + np.uint64(0x3F000000)).astype(np.uint32)


@pytest.mark.parametrize("rng", ["philox", "minstd"])
def test_products_shape_full_batches(c2, rng):
    import legion_b200 as L
    from oracle import oracle as O
    d, cfg = c2
    B, fanout = cfg["batch"], cfg["fanout"]
    mode = L.RNG_PHILOX if rng == "philox" else L.RNG_MINSTD
    r = L.Runner(d.n_nodes, d.dim, B, fanout, rng_mode=mode, rng_seed=42, enable_hotness=True, n_lanes=4)
    ipd, ixd = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices)
    r.bind_topology(ipd, ixd)
    feats_d = L.DevArray.from_numpy(d.features)
    r.bind_features(feats_d)
    train = d.train_ids
    r.bind_seeds(L.MODE_TRAIN, L.DevArray.from_numpy(train), L.DevArray.from_numpy(d.labels[train]))
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=mode, rng_seed=42, n_threads=8)
    smp.enable_hotness()
    # presampling over 4 lanes in flight, then the planner: everything cached on one GPU
    n_pre = 6
    for step in range(n_pre):
        r.batch_generate(L.MODE_TRAIN, B, step, pipe=step % 4)
        r.run_batch(with_features=False, is_presc=True)
        smp.sample(train[step * B:(step + 1) * B], step=step)
    nh, th = r.hotness()
    assert np.array_equal(nh.numpy(), smp.node_hotness) and np.array_equal(th.numpy(), smp.topo_hotness)
    assert int(th.numpy().sum()) == r_totals(r)[1]
    order = L.hot_order(nh)
    assert np.array_equal(order.numpy(), O.hot_order(smp.node_hotness))
    slot = L.place(order, d.n_nodes, 1)
    shard = L.fill_feature_shard(order, d.n_nodes, 1, 0, feats_d, d.dim)
    r.bind_feature_cache([shard], slot, d.n_nodes)
    for step in range(3):
        seeds = train[step * B:(step + 1) * B]
        r.batch_generate(L.MODE_TRAIN, B, step, pipe=step % 4)
        r.run_batch(with_features=True)
        got = r.fetch()
        want = smp.sample(seeds, step=step)
        nc, ec = got["nc"], got["ec"]
        total, n_e = int(nc[0]), int(ec[0])
        for k in ("nc", "ec"):
            assert np.array_equal(got[k], want[k]), (rng, step, k)
        assert np.array_equal(got["sampled_ids"], want["sampled_ids"][:total])
        for k in ("agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off"):
            assert np.array_equal(got[k], want[k][:n_e]), (rng, step, k)
        # size-independent properties
        ids = got["sampled_ids"]
        assert len(np.unique(ids)) == total                                           # dedup
        assert np.array_equal(ids[got["agg_src_off"]], got["agg_src_ids"])            # relabelling round trip
        assert np.array_equal(ids[got["agg_dst_off"]], got["agg_dst_ids"])
        assert nc[9] == nc[4] + nc[6] + nc[8] and ec[4] == n_e and nc[2] == ec[4] - ec[3]
        assert got["agg_dst_off"][:ec[3]].max() < nc[5] and got["agg_dst_off"].max() < nc[7]   # dst nodes are a prefix of src nodes
        assert np.array_equal(got["features"].view(np.uint32), _closed_form_features(ids, d.dim))
        assert np.array_equal(got["labels"], d.labels[seeds])
        deg = np.diff(d.indptr)
        if rng == "philox":        # exact neighbourhood whenever the fanout covers the degree
            small = np.flatnonzero(deg[seeds] <= fanout[0])[:200]
            e1 = int(ec[3])
            for i in small:
                nb = d.indices[d.indptr[seeds[i]]:d.indptr[seeds[i] + 1]]
                assert np.array_equal(np.sort(got["agg_src_ids"][:e1][got["agg_dst_off"][:e1] == i]), np.sort(nb))
        # idempotence: the same step again in another lane gives the same bytes
        r.batch_generate(L.MODE_TRAIN, B, step, pipe=(step + 1) % 4)
        r.run_batch(with_features=True)
        again = r.fetch()
        for k in ("nc", "ec", "sampled_ids", "agg_src_off", "agg_dst_off", "features"):
            assert np.array_equal(got[k], again[k]), (rng, step, k)
    assert r.status() == 0
    assert r.tier_counts()[2] == 0          # everything was a cache hit
    r.close()


def r_totals(r):
    import ctypes as C
    import legion_b200 as L
    out = (C.c_int64 * 2)()
    L._lib.check(L.lib().lgn_sampling_totals(r.handle, None, out, 0), "lgn_sampling_totals")
    return int(out[0]), int(out[1])
