"""Parity tests proper: the CUDA hot path (through the C-ABI) against the CPU oracle on the
same seeded inputs.  Integer / index / byte work throughout => every comparison is bit-exact.
Edge cases follow the reference's own branches: padded (-1) seeds (Kernels.cu:81-83,385),
clamped last batch (Kernels.cu:224), zero-degree nodes (Kernels.cu:399), duplicates in the
frontier (Kernels.cu:371-373), cache hit / miss tiers (Kernels.cu:692-699)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INT_KEYS = ("nc", "ec", "sampled_ids", "agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off")


def _oracle_batch(O, smp, seeds, step, epoch=0):
    out = smp.sample(seeds, step=step, epoch=epoch)
    total, n_e = int(out["nc"][0]), int(out["ec"][0])
    res = {k: out[k] for k in ("nc", "ec")}
    res["sampled_ids"] = out["sampled_ids"][:total]
    for k in ("agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off"):
        res[k] = out[k][:n_e]
    return res


def _assert_same(got, want, keys=INT_KEYS, ctx=""):
    for k in keys:
        assert got[k].shape == want[k].shape, f"{ctx}{k}: shape {got[k].shape} vs {want[k].shape}"
        assert np.array_equal(got[k], want[k]), f"{ctx}{k} differs (first at {np.flatnonzero(got[k] != want[k])[:5]})"


def _make_runner(L, d, batch, fanout, rng_mode, seed=7, feat=True, **kw):
    r = L.Runner(d.n_nodes, d.dim if feat else 0, batch, fanout, rng_mode=rng_mode, rng_seed=seed, **kw)
    ip, ix = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices)
    r.bind_topology(ip, ix)
    if feat:
        r.bind_features(L.DevArray.from_numpy(d.features))
    return r


@pytest.mark.parametrize("dedup", ["hash", "direct"])
@pytest.mark.parametrize("rng", ["minstd", "philox"])
@pytest.mark.parametrize("fanout", [[25, 10], [15, 10, 5], [3], [40, 2]])
def test_sampling_bit_exact(c1, rng, fanout, dedup, monkeypatch):
    import legion_b200 as L
    from oracle import oracle as O
    monkeypatch.setenv("LGN_DEDUP", dedup)     # batch-sized hash table vs direct int32[N] map: same bytes out
    mode = L.RNG_MINSTD if rng == "minstd" else L.RNG_PHILOX
    B = 1024
    r = _make_runner(L, c1, B, fanout, mode, feat=False)
    smp = O.Sampler(c1.indptr, c1.indices, fanout, rng_mode=mode, rng_seed=7)
    train = c1.train_ids
    for step in range(3):
        seeds = train[step * B:(step + 1) * B]
        r.batch_from_host(seeds, None, step=step, pipe=step % 2)
        for h in range(len(fanout)):
            r.sample_hop(h)
        r.finish_batch()
        got = r.fetch(with_features=False)
        want = _oracle_batch(O, smp, seeds, step)
        _assert_same(got, want, ctx=f"{rng} {fanout} step {step}: ")
        assert r.status() == 0
    r.close()


def test_philox_epoch_and_step_offset_key_the_stream(c1):
    """Philox counter = (slot, epoch, hop, step): the same seeds redrawn in another epoch / at another step offset give
    other neighbourhoods, each bit-exact against the oracle keyed the same way (the reference's
    minstd stream redraws identical neighbourhoods every epoch, Kernels.cu:402-405)."""
    import legion_b200 as L
    from oracle import oracle as O
    B, fanout = 512, [10, 5]
    r = _make_runner(L, c1, B, fanout, L.RNG_PHILOX, feat=False)
    smp = O.Sampler(c1.indptr, c1.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=7)
    seeds = c1.train_ids[:B]
    seen = []
    for epoch, off, step in ((0, 0, 3), (1, 0, 3), (2, 0, 3), (1, 100, 3), (1, 0, 103)):
        r.set_epoch(epoch, off)
        r.batch_from_host(seeds, None, step=step)
        for h in range(len(fanout)):
            r.sample_hop(h)
        r.finish_batch()
        got = r.fetch(with_features=False)
        _assert_same(got, _oracle_batch(O, smp, seeds, step + off, epoch=epoch), ctx=f"epoch {epoch} off {off}: ")
        seen.append(got["agg_src_ids"].copy())
    assert not np.array_equal(seen[0], seen[1]) and not np.array_equal(seen[1], seen[2])
    assert np.array_equal(seen[3], seen[4])              # offset + step is what enters the counter
    r.close()


@pytest.mark.parametrize("dedup", ["hash", "direct"])
def test_sampling_edge_cases(small, dedup, monkeypatch):
    """ragged / empty / padded / duplicate seeds, zero-degree nodes."""
    import legion_b200 as L
    from oracle import oracle as O
    monkeypatch.setenv("LGN_DEDUP", dedup)
    d = small
    # make some isolated nodes and a hub so deg==0, deg<f, deg>f all occur
    indptr, indices = d.indptr.copy(), d.indices.copy()
    fanout = [5, 4]
    for mode in (L.RNG_MINSTD, L.RNG_PHILOX):
        r = L.Runner(d.n_nodes, 0, 64, fanout, rng_mode=mode, rng_seed=123)
        r.bind_topology(L.DevArray.from_numpy(indptr), L.DevArray.from_numpy(indices))
        smp = O.Sampler(indptr, indices, fanout, rng_mode=mode, rng_seed=123)
        cases = [
            np.array([], np.int32),
            np.array([5], np.int32),
            np.array([7, -1, 9, -1], np.int32),                        # padding ids (Kernels.cu:81-83)
            np.array([11, 11, 12, 11, 40, 12], np.int32),              # duplicate seeds (lp_sage layout)
            np.arange(0, 64, dtype=np.int32),
            np.full(10, -1, np.int32),
        ]
        for step, seeds in enumerate(cases):
            r.batch_from_host(seeds, None, step=step)
            for h in range(len(fanout)):
                r.sample_hop(h)
            r.finish_batch()
            got = r.fetch(with_features=False)
            want = _oracle_batch(O, smp, seeds, step)
            _assert_same(got, want, ctx=f"mode {mode} case {step}: ")
        r.close()


def test_full_neighbourhood_when_fanout_covers_degree():
    """north star: exact full-neighbourhood results when fanout >= degree (philox mode)."""
    import legion_b200 as L
    d = L.synth.make_dataset(5_000, 8.0, 4, kmax=3, with_features=False)
    deg = np.diff(d.indptr)
    f = int(deg.max())
    assert f <= 256
    seeds = np.arange(0, 200, dtype=np.int32)
    r = L.Runner(d.n_nodes, 0, 256, [f], rng_mode=L.RNG_PHILOX, rng_seed=1)
    r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
    r.batch_from_host(seeds, None, step=0)
    r.sample_hop(0)
    r.finish_batch()
    got = r.fetch(with_features=False)
    src_of_edge, dst_of_edge = got["agg_dst_ids"], got["agg_src_ids"]
    pos = 0
    for s in seeds:
        nb = d.indices[d.indptr[s]:d.indptr[s + 1]]
        assert np.array_equal(dst_of_edge[pos:pos + len(nb)], nb)
        assert np.all(src_of_edge[pos:pos + len(nb)] == s)
        pos += len(nb)
    assert pos == len(dst_of_edge)
    r.close()


def test_batch_generate_matches_reference_indexing(small):
    """op 0 incl. the clamped last batch (Kernels.cu:224-227) and labels."""
    import legion_b200 as L
    from oracle import oracle as O
    d = small
    ids = d.test_ids.astype(np.int32)
    labels = d.labels[ids]
    r = L.Runner(d.n_nodes, 0, 16, [2])
    r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
    r.bind_seeds(L.MODE_TEST, L.DevArray.from_numpy(ids), L.DevArray.from_numpy(labels))
    B = 16
    steps = (len(ids) - 1) // B + 1
    for counter in range(steps):
        r.batch_generate(L.MODE_TEST, B, counter)
        r.sample_hop(0)
        r.finish_batch()
        got = r.fetch(with_features=False)
        w_ids, w_lab = O.batch_generate(ids, labels, B, counter)
        assert np.array_equal(got["sampled_ids"][:len(w_ids)], w_ids)
        assert np.array_equal(got["labels"], w_lab)
        assert got["nc"][4] == len(w_ids)
    r.close()


@pytest.mark.parametrize("gather", ["bulk", "ldg"])
@pytest.mark.parametrize("dim", [100, 128, 256, 7])
@pytest.mark.parametrize("kg,frac,host", [(1, 1.0, False), (1, 0.3, True), (4, 0.5, True), (8, 1.0, False), (0, 0.0, True)])
def test_gather_bit_exact(dim, kg, frac, host, gather, monkeypatch):
    """all three tiers: local shard, 'peer' shards (separate allocations addressed through the
    shard table, emulated on one GPU), base matrix in mapped host memory (UVA zero-copy)."""
    import legion_b200 as L
    from oracle import oracle as O
    monkeypatch.setenv("LGN_GATHER", gather)   # cp.async.bulk (TMA) thread-per-row vs 128-bit LDG warp-per-row
    d = L.synth.make_dataset(20_000, 10.0, dim, n_class=5)
    fanout = [10, 5]
    B = 512
    r = L.Runner(d.n_nodes, dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=3, part=0)
    r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
    base = L.MappedHostArray.from_numpy(d.features) if host else L.DevArray.from_numpy(d.features)
    r.bind_features(base)
    slot_h, shards_h, cap = None, [], 1
    if kg > 0:
        rng = np.random.default_rng(5)
        counts = rng.integers(0, 50, d.n_nodes).astype(np.uint32)
        order_h = O.hot_order(counts)
        cap = int(np.ceil(d.n_nodes * frac / kg))
        slot_h = O.place(order_h, cap, kg)
        shards_h = [O.fill_feature_shard(order_h, cap, kg, j, d.features) for j in range(kg)]
        order_d = L.hot_order(L.DevArray.from_numpy(counts))
        assert np.array_equal(order_d.numpy(), order_h)
        slot_d = L.place(order_d, cap, kg)
        assert np.array_equal(slot_d.numpy(), slot_h)
        shards_d = [L.fill_feature_shard(order_d, cap, kg, j, base, dim) for j in range(kg)]
        for j in range(kg):   # cache contents bit-exact
            assert np.array_equal(shards_d[j].numpy().view(np.uint32), shards_h[j].view(np.uint32))
        r.bind_feature_cache(shards_d, slot_d, cap)
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=3)
    for step in range(2):
        seeds = d.train_ids[step * B:(step + 1) * B]
        r.batch_from_host(seeds, d.labels[seeds], step=step)
        r.run_batch(with_features=True)
        got = r.fetch()
        want = _oracle_batch(O, smp, seeds, step)
        _assert_same(got, want)
        total = int(want["nc"][0])
        ref = np.zeros((total, dim), np.float32)
        tiers = O.gather(want["sampled_ids"], 0, total, slot_h, cap, shards_h, d.features, ref, tiers=True)
        assert np.array_equal(got["features"].view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(got["labels"], d.labels[seeds])
        tc = r.tier_counts(reset=True)
        assert tc[0] + tc[1] + tc[2] == total
        if kg > 0:
            assert tc[0] == tiers[0] and tc[1] == tiers[1:kg].sum() and tc[2] == tiers[kg]
    r.close()


@pytest.mark.parametrize("kg,frac_repl,frac_shard", [(4, 0.2, 0.5), (2, 1.0, 0.0), (8, 0.0, 1.0), (4, 0.1, 0.2)])
def test_hybrid_placement_gather(kg, frac_repl, frac_shard):
    """B200 extension: hottest ranks replicated on every GPU, warm ranks partitioned, cold ranks on the host tier.
    Every emulated GPU of the clique (its own slot map + shard, peers addressed through the shard table) must
    gather the same bit-exact rows; replicated rows must be served locally."""
    import legion_b200 as L
    from oracle import oracle as O
    dim = 100
    d = L.synth.make_dataset(20_000, 10.0, dim, n_class=5)
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 50, d.n_nodes).astype(np.uint32)
    order_h = O.hot_order(counts)
    n_repl = int(d.n_nodes * frac_repl)
    cap = n_repl + int(np.ceil(d.n_nodes * frac_shard / kg))
    base = L.MappedHostArray.from_numpy(d.features)
    order_d = L.hot_order(L.DevArray.from_numpy(counts))
    shards_h = [O.fill_feature_shard_hybrid(order_h, cap, kg, j, n_repl, d.features) for j in range(kg)]
    shards_d = [L.fill_feature_shard_hybrid(order_d, cap, kg, j, n_repl, base, dim) for j in range(kg)]
    for j in range(kg):
        assert np.array_equal(shards_d[j].numpy().view(np.uint32), shards_h[j].view(np.uint32))
    fanout, B = [10, 5], 512
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=3)
    seeds = d.train_ids[:B]
    want = smp.sample(seeds, step=0)
    total = int(want["nc"][0])
    for me in range(0, kg, max(1, kg // 2)):
        slot_h = O.place_hybrid(order_h, cap, kg, n_repl, me)
        slot_d = L.place_hybrid(order_d, cap, kg, n_repl, me)
        assert np.array_equal(slot_d.numpy(), slot_h)
        r = L.Runner(d.n_nodes, dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=3, part=me)
        r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
        r.bind_features(base)
        r.bind_feature_cache(shards_d, slot_d, cap)
        r.batch_from_host(seeds, None, step=0)
        r.run_batch(with_features=True)
        got = r.fetch()
        assert np.array_equal(got["features"].view(np.uint32), d.features[want["sampled_ids"][:total]].view(np.uint32))
        ref = np.zeros((total, dim), np.float32)
        tiers = O.gather(want["sampled_ids"], 0, total, slot_h, cap, shards_h, d.features, ref, tiers=True)
        tc = r.tier_counts()
        assert tc == [int(tiers[me]), int(tiers[:kg].sum() - tiers[me]), int(tiers[kg])]
        hot = np.isin(want["sampled_ids"][:total], order_h[:n_repl]).sum()
        assert tc[0] >= hot                                                  # every replicated row was a local hit
        r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gather", ["bulk", "ldg"])
@pytest.mark.parametrize("dim,kg,frac_repl,frac_part", [(128, 4, 0.2, 0.5), (100, 1, 1.0, 0.0), (256, 8, 0.0, 1.0), (7, 2, 0.1, 0.2), (128, 2, 0.0, 0.0)])
def test_compact_placement_gather(dim, kg, frac_repl, frac_part, gather, monkeypatch):
    """compact (L2-resident) placement map: the map words, every shard and the gathered rows + tier counts of every emulated
    GPU equal the numpy restatement; n_repl = N bound without a map (rows addressed by node id) gathers the same rows."""
    import legion_b200 as L
    from oracle import oracle as O
    monkeypatch.setenv("LGN_GATHER", gather)
    N = 20_011                                  # not a multiple of the 96-node record
    d = L.synth.make_dataset(N, 10.0, dim, n_class=5)
    rng = np.random.default_rng(2)
    counts = rng.integers(0, 50, N).astype(np.uint32)
    order_h = O.hot_order(counts)
    order_d = L.hot_order(L.DevArray.from_numpy(counts))
    n_repl = int(N * frac_repl)
    n_part = min(N - n_repl, int(N * frac_part))
    cap = max(1, n_repl + (n_part + kg - 1) // kg)
    base = L.MappedHostArray.from_numpy(d.features)
    cmap_d = L.place_compact(order_d, n_repl, n_part)
    shards_d = [L.fill_feature_shard_compact(cmap_d, N, n_repl, kg, j, base, dim, cap) for j in range(kg)]
    fanout, B = [10, 5], 512
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=3)
    seeds = d.train_ids[:B]
    want = _oracle_batch(O, smp, seeds, 0)
    total = int(want["nc"][0])
    shards_h = None
    for me in range(0, kg, max(1, kg // 2)):
        words_h, slot_h = O.place_compact(order_h, n_repl, n_part, kg, me, cap)
        assert np.array_equal(cmap_d.numpy(), words_h)
        if shards_h is None:
            shards_h = []
            for j in range(kg):
                _, slot_j = O.place_compact(order_h, n_repl, n_part, kg, j, cap)
                shards_h.append(O.fill_feature_shard_compact(slot_j, cap, j, d.features))
                assert np.array_equal(shards_d[j].numpy().view(np.uint32), shards_h[j].view(np.uint32))
        r = L.Runner(N, dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=3, part=me)
        r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
        r.bind_features(base)
        r.bind_feature_cache_compact(shards_d, cmap_d, n_repl, cap)
        r.batch_from_host(seeds, None, step=0)
        r.run_batch(with_features=True)
        got = r.fetch()
        _assert_same(got, want)
        assert np.array_equal(got["features"].view(np.uint32), d.features[want["sampled_ids"][:total]].view(np.uint32))
        ref = np.zeros((total, dim), np.float32)
        tiers = O.gather(want["sampled_ids"], 0, total, slot_h, cap, shards_h, d.features, ref, tiers=True)
        assert np.array_equal(ref.view(np.uint32), got["features"].view(np.uint32))
        assert r.tier_counts() == [int(tiers[me]), int(tiers[:kg].sum() - tiers[me]), int(tiers[kg])]
        r.close()
    if n_repl == N:      # the whole matrix resident in id order: no map
        r = L.Runner(N, dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=3, part=0)
        r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
        r.bind_features(base)
        resident = L.DevArray.from_numpy(d.features)
        assert np.array_equal(shards_d[0].numpy().view(np.uint32), d.features.view(np.uint32))    # id order IS the matrix
        r.bind_feature_cache_compact([resident], None, N, N)
        r.batch_from_host(seeds, None, step=0)
        r.run_batch(with_features=True)
        got = r.fetch()
        assert np.array_equal(got["features"].view(np.uint32), d.features[want["sampled_ids"][:total]].view(np.uint32))
        assert r.tier_counts() == [total, 0, 0]
        r.close()


@pytest.mark.parametrize("kg,budget_frac", [(1, 0.3), (2, 0.4), (8, 0.2), (8, 0.6), (4, 2.0)])
def test_plan_hybrid_picks_the_cheapest_split(kg, budget_frac):
    """B200 placement model: of the 101 candidate splits (replicated / partitioned / host) the one with the smallest
    expected gather time under the three tier bandwidths is chosen; restated here with numpy on the same candidates."""
    import legion_b200 as L
    n, dim = 50_000, 64
    rng = np.random.default_rng(5)
    counts = np.sort((rng.pareto(1.2, n) * 20).astype(np.uint32))[::-1].copy()       # hot order, heavy tail
    row = dim * 4
    budget = int(n * row * budget_frac)
    bw_l, bw_p, bw_h = 3272.0, 640.0, 50.0
    prior = 0.5 if kg == 8 else 0.0
    n_repl, cap, cost = L.plan_hybrid(L.DevArray.from_numpy(counts), dim, budget, kg, bw_l, bw_p, bw_h, prior=prior)
    H = np.concatenate([[0], np.cumsum(counts.astype(np.uint64))]).astype(np.float64) + prior * np.arange(n + 1)
    budget_rows = budget // row
    best = None
    for i in range(101):
        rp = 0 if kg == 1 else min(budget_rows, n) * i // 100
        part = max(0, min(n - rp, (budget_rows - rp) * kg))
        t = H[rp] / bw_l + (H[rp + part] - H[rp]) * (1 / kg / bw_l + (kg - 1) / kg / bw_p) + (H[n] - H[rp + part]) / bw_h
        if best is None or t < best[0]:
            best = (t, rp, rp + (part + kg - 1) // kg)
        if kg == 1:
            break
    assert (n_repl, cap) == (best[1], max(1, best[2]))
    assert abs(cost - best[0] * row) <= 1e-9 * abs(cost)
    if budget_frac >= 1.0:
        assert H[n_repl] == H[n]                # everything fits on every GPU: all the presampled hotness is served locally
    assert cap * row <= budget + row


def test_empty_batch_is_accepted(small):
    """a partition without valid/test ids yields batch size 0 (lgn_coordinate); the reference runs an empty batch."""
    import legion_b200 as L
    d = small
    r = L.Runner(d.n_nodes, 0, 16, [3, 2])
    r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
    ids = d.valid_ids[:5].astype(np.int32)
    r.bind_seeds(L.MODE_VALID, L.DevArray.from_numpy(ids), L.DevArray.from_numpy(d.labels[ids]))
    r.batch_generate(L.MODE_VALID, 0, 0)
    r.run_batch(with_features=False)
    got = r.fetch(with_features=False)
    assert int(got["nc"][0]) == 0 and int(got["ec"][0]) == 0 and r.status() == 0
    r.close()


def test_debug_shard_read_copies_shard_rows():
    """lgn_debug_shard_read (diagnostic entry point): every output row must be a bit-exact copy of SOME row of the
    bound shards, peers_only must leave this GPU's own shard out, and argument checks must hold."""
    import legion_b200 as L
    dim, kg, cap = 64, 4, 500
    rng = np.random.default_rng(5)
    shards_h = [(rng.random((cap, dim), dtype=np.float32) * 0.5 + j).astype(np.float32) for j in range(kg)]   # shard j holds values in [j, j+0.5]
    shards_d = [L.DevArray.from_numpy(x) for x in shards_h]
    n_nodes = cap * kg
    slot = L.DevArray.from_numpy(np.arange(n_nodes, dtype=np.int32))
    base = L.DevArray.from_numpy(np.zeros((n_nodes, dim), np.float32))
    r = L.Runner(n_nodes, dim, 256, [4], part=1)
    r.bind_features(base)
    r.bind_feature_cache(shards_d, slot, cap)
    n_rows = 1000
    for peers_only in (False, True):
        ms = r.debug_shard_read(n_rows, cap, peers_only=peers_only, repeats=2)
        assert ms > 0
        out = L.DevArray((n_rows, dim), np.float32, ptr=r.view(0).features, owner=False).numpy()
        owner = np.floor(out[:, 0]).astype(int)
        assert ((owner >= 0) & (owner < kg)).all()
        if peers_only:
            assert (owner != 1).all()
        table = {j: {row.tobytes() for row in shards_h[j]} for j in range(kg)}
        assert all(out[i].tobytes() in table[owner[i]] for i in range(0, n_rows, 7))
    with pytest.raises(L._lib.LegionError):
        r.debug_shard_read(n_rows, cap + 1)
    r.close()


def test_shared_allocation_round_trip():
    """lgn_shared_alloc / lgn_shared_import / lgn_shared_free (VMM shards shared as file descriptors): a second mapping
    of the same allocation, imported through the descriptor, must see the bytes written through the first one, and a
    gather must read it like any other shard."""
    import legion_b200 as L
    rows, dim = 3000, 32
    a, fd, mapped = L.shared_alloc((rows, dim), np.float32)
    assert fd >= 0 and mapped >= rows * dim * 4
    want = np.random.default_rng(2).random((rows, dim), dtype=np.float32)
    import ctypes as C
    L._lib.check(L.lib().lgn_copy_h2d(C.c_void_p(a.ptr), want.ctypes.data_as(C.c_void_p), C.c_int64(want.nbytes)), "h2d")
    b = L.shared_import(fd, mapped, (rows, dim), np.float32)
    os.close(fd)
    assert b.ptr != a.ptr
    assert np.array_equal(b.numpy().view(np.uint32), want.view(np.uint32))
    # the imported mapping as a cache shard
    r = L.Runner(rows, dim, 128, [3])
    r.bind_features(L.DevArray.from_numpy(np.zeros((rows, dim), np.float32)))
    r.bind_feature_cache([b], L.DevArray.from_numpy(np.arange(rows, dtype=np.int32)), rows)
    ms = r.debug_shard_read(400, rows, peers_only=False, repeats=1)
    assert ms > 0
    r.close()
    L.shared_free(b)
    L.shared_free(a)
    with pytest.raises(L._lib.LegionError):
        L.shared_free(a)


def test_presampling_hotness_and_planner(c1):
    """presampling epoch: node / topology hotness, max ids, hot order, shards, cost model."""
    import legion_b200 as L
    from oracle import oracle as O
    d = c1
    fanout, B, steps = [25, 10], 1024, 5
    for mode in (L.RNG_MINSTD, L.RNG_PHILOX):
        r = L.Runner(d.n_nodes, d.dim, B, fanout, rng_mode=mode, rng_seed=11, enable_hotness=True)
        ipd, ixd = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices)
        r.bind_topology(ipd, ixd)
        ids = d.train_ids
        r.bind_seeds(L.MODE_TRAIN, L.DevArray.from_numpy(ids), L.DevArray.from_numpy(d.labels[ids]))
        smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=mode, rng_seed=11)
        smp.enable_hotness()
        max_ids = 0
        for step in range(steps):
            r.batch_generate(L.MODE_TRAIN, B, step, pipe=step % 2)
            r.run_batch(with_features=False, is_presc=True)
            w = smp.sample(ids[step * B:(step + 1) * B], step=step)
            max_ids = max(max_ids, int(w["nc"][0]))
        nh, th = r.hotness()
        assert np.array_equal(nh.numpy(), smp.node_hotness)
        assert np.array_equal(th.numpy(), smp.topo_hotness)
        assert r.max_ids() == max_ids
        # candidate selection
        qf, af = L.hot_order(nh, want_sorted=True)
        qt, at = L.hot_order(th, want_sorted=True)
        qf_h, qt_h = O.hot_order(smp.node_hotness), O.hot_order(smp.topo_hotness)
        assert np.array_equal(qf.numpy(), qf_h) and np.array_equal(qt.numpy(), qt_h)
        assert np.array_equal(af.numpy(), smp.node_hotness[qf_h])
        # cost model (restricted cache so both tiers compete)
        cache_mem = 20_000_000
        for kg in (1, 2):
            got = L.cost_model(af, at, qt, ipd, d.dim, cache_mem, kg, 123456, [max_ids] * kg, steps)
            want = O.cost_model(smp.node_hotness[qf_h], smp.topo_hotness[qt_h], qt_h, d.indptr, d.dim, cache_mem, kg,
                                123456, [max_ids] * kg, steps)
            assert got == want[:2], (got, want)
        # topology shards
        kg, cap = 4, 9000
        for j in range(kg):
            ip_d, ix_d, n = L.fill_topo_shard(qt, cap, kg, j, ipd, ixd)
            ip_h, ix_h = O.fill_topo_shard(qt_h, cap, kg, j, d.indptr, d.indices)
            assert n == len(ix_h)
            assert np.array_equal(ip_d.numpy(), ip_h)
            assert np.array_equal(ix_d.numpy(n), ix_h)
        r.close()


def test_topology_cache_tiers(c1):
    """train-time sampler over a sharded topology cache: hits read shard CSRs, misses the
    base CSR in mapped host memory -- same samples as the flat CSR (Kernels.cu:389-410)."""
    import legion_b200 as L
    from oracle import oracle as O
    d = c1
    fanout, B = [25, 10], 1024
    rng = np.random.default_rng(9)
    counts = rng.integers(0, 100, d.n_nodes).astype(np.uint32)
    ipd, ixd = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices)
    qt = L.hot_order(L.DevArray.from_numpy(counts))
    kg, cap = 4, 15_000    # 60 % of the nodes cached over 4 shards
    slot = L.place(qt, cap, kg)
    shards = [L.fill_topo_shard(qt, cap, kg, j, ipd, ixd) for j in range(kg)]
    for mode in (L.RNG_MINSTD, L.RNG_PHILOX):
        r = L.Runner(d.n_nodes, 0, B, fanout, rng_mode=mode, rng_seed=2)
        r.bind_topology(L.MappedHostArray.from_numpy(d.indptr), L.MappedHostArray.from_numpy(d.indices))
        r.bind_topology_cache([s[0] for s in shards], [s[1] for s in shards], slot, cap)
        smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=mode, rng_seed=2)
        for step in range(2):
            seeds = d.train_ids[step * B:(step + 1) * B]
            r.batch_from_host(seeds, None, step=step)
            r.run_batch(with_features=False)
            _assert_same(r.fetch(with_features=False), _oracle_batch(O, smp, seeds, step))
        r.close()


def test_run_batch_overlap_equals_stepwise_and_pipes(c1):
    """two-stream RunOnce DAG == operator-by-operator execution; alternating pipe slots keep
    the previous batch intact (double buffering, Server.cu:325-327)."""
    import legion_b200 as L
    d = c1
    fanout, B = [25, 10], 1024
    r = _make_runner(L, d, B, fanout, L.RNG_PHILOX)
    seeds0, seeds1 = d.train_ids[:B], d.train_ids[B:2 * B]
    r.batch_from_host(seeds0, d.labels[seeds0], step=0, pipe=0)
    r.run_batch(with_features=True)
    a0 = r.fetch()
    r.batch_from_host(seeds1, d.labels[seeds1], step=1, pipe=1)
    for seg in range(len(fanout) + 1):
        if seg > 0:
            r.sample_hop(seg - 1)
        r.gather_segment(seg)
    r.finish_batch()
    b1 = r.fetch()
    r.pipe = 0
    a0_again = r.fetch()
    per_pipe = ("nc", "ec", "sampled_ids", "agg_src_off", "agg_dst_off", "features", "labels")   # raw-id edge lists are shared scratch
    for k in per_pipe:
        assert np.array_equal(a0[k], a0_again[k]), k
    # and the same batch through both paths
    r.batch_from_host(seeds1, d.labels[seeds1], step=1, pipe=0)
    r.run_batch(with_features=True)
    b1_dag = r.fetch()
    for k in INT_KEYS + ("features", "labels"):
        assert np.array_equal(b1[k], b1_dag[k]), k
    r.close()


def test_hash_dedup_shrunk_table_and_overflow(c1, monkeypatch):
    """the hash table sized from presampling (2.5 x max unique ids) gives the same bytes; a table that is too
    small reports LGN_E_CAPACITY instead of corrupting the batch silently."""
    import legion_b200 as L
    from oracle import oracle as O
    monkeypatch.setenv("LGN_DEDUP", "hash")
    d, fanout, B = c1, [25, 10], 1024
    r = _make_runner(L, d, B, fanout, L.RNG_PHILOX, feat=False)
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=7)
    seeds = d.train_ids[:B]
    want = _oracle_batch(O, smp, seeds, 0)
    r.set_dedup_capacity(int(want["nc"][0]))
    for step in range(2):
        r.batch_from_host(seeds, None, step=0, pipe=step)
        r.run_batch(with_features=False)
        _assert_same(r.fetch(with_features=False), want)
    assert r.status() == 0
    r.set_dedup_capacity(64)                       # far too small for ~20k unique ids
    r.batch_from_host(seeds, None, step=0, pipe=0)
    r.run_batch(with_features=False)
    r.read_counters()
    assert r.status() == L._lib.E_CAPACITY
    r.close()


def test_cuda_graph_replay_matches_eager(c1):
    """on a real stream lgn_run_batch replays a captured CUDA graph from the second call on: same bytes as the
    eager path and as the oracle, across lanes, after re-binding (graphs are invalidated) and with hotness."""
    import legion_b200 as L
    from oracle import oracle as O
    d, fanout, B = c1, [25, 10], 1024
    r = _make_runner(L, d, B, fanout, L.RNG_PHILOX, n_lanes=3, enable_hotness=True)
    streams = [L.Stream() for _ in range(3)]
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=7)
    smp.enable_hotness()
    for step in range(9):                       # 3 calls per lane: eager, capture + replay, replay
        q = step % 3
        seeds = d.train_ids[step * B:(step + 1) * B]
        presc = step >= 6
        r.batch_from_host(seeds, d.labels[seeds], step=step, stream=streams[q].handle, pipe=q)
        r.run_batch(with_features=not presc, is_presc=presc, stream=streams[q].handle)
        got = r.fetch(with_features=not presc, stream=streams[q].handle)
        smp.topo_hotness_enabled = presc
        if presc:
            want = _oracle_batch(O, smp, seeds, step)
        else:
            keep = (smp.topo_hotness, smp.node_hotness)
            smp.topo_hotness = smp.node_hotness = None
            want = _oracle_batch(O, smp, seeds, step)
            smp.topo_hotness, smp.node_hotness = keep
        _assert_same(got, want, ctx=f"step {step}: ")
        if not presc:
            assert np.array_equal(got["features"].view(np.uint32), d.features[want["sampled_ids"]].view(np.uint32))
    nh, th = r.hotness()
    assert np.array_equal(nh.numpy(), smp.node_hotness) and np.array_equal(th.numpy(), smp.topo_hotness)
    r.bind_features(L.DevArray.from_numpy(d.features * 2.0))       # re-binding must not replay a stale graph
    seeds = d.train_ids[:B]
    r.batch_from_host(seeds, None, step=0, stream=streams[0].handle, pipe=0)
    r.run_batch(with_features=True, stream=streams[0].handle)
    got = r.fetch(stream=streams[0].handle)
    assert np.array_equal(got["features"], d.features[got["sampled_ids"]] * 2.0)
    assert r.status() == 0
    r.close()
    for s in streams:
        s.close()


def test_capacity_overflow_is_reported(small):
    """reference: silent overflow of the 1.2x feature buffer (Server.cu:275); here: status code."""
    import legion_b200 as L
    d = small
    r = L.Runner(d.n_nodes, d.dim, 256, [10, 5], max_feature_rows=300)
    r.bind_topology(L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices))
    r.bind_features(L.DevArray.from_numpy(d.features))
    r.batch_from_host(d.train_ids[:256], None, step=0)
    r.run_batch(with_features=True)
    nc, _ = r.read_counters()
    assert nc[0] > 300
    assert r.status() == L._lib.E_CAPACITY
    with pytest.raises(L.LegionError):
        r.batch_from_host(np.zeros(1000, np.int32), None)
    r.close()


def test_parity_suite_against_the_assert_build():
    """compute-sanitizer is closed on the measurement pool: the sampling / gather parity cases are repeated against the
    debug build of the library (-DLGN_DEBUG: a device assert in front of every indexed write, common.cuh).  A tripped
    assert surfaces as a CUDA error and fails the inner run."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dbg = os.path.join(root, "legion-1_b200", "_build", "liblegion_b200_dbg.so")
    if not os.path.exists(dbg):
        pytest.skip("debug build missing (make -C legion-1_b200/csrc debug)")
    if os.environ.get("LGN_LIBRARY"):
        pytest.skip("already running against an explicit library")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                          "sampling_bit_exact or edge_cases or full_neighbourhood or epoch or presampling or gather_bit_exact or compact_placement or run_batch"],
                         env=dict(os.environ, LGN_LIBRARY=dbg), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout


def test_slot_status_and_attached_buffers(small):
    """lgn_sync_pipe_status reports a slot's overflow (the server dies on it before IPCPost); lgn_attach_buffers makes the
    kernels write the caller's buffers and rejects buffers that cannot even hold the seeds or are misaligned."""
    import ctypes as C
    import legion_b200 as L
    from oracle import oracle as O
    d = small
    fanout, B = [5, 4], 64
    ip, ix, ft = L.DevArray.from_numpy(d.indptr), L.DevArray.from_numpy(d.indices), L.DevArray.from_numpy(d.features)
    r = L.Runner(d.n_nodes, d.dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=9, max_feature_rows=10)   # feature buffer too small
    r.bind_topology(ip, ix); r.bind_features(ft)
    seeds = d.train_ids[:B]
    r.batch_from_host(seeds, None, step=0)
    r.run_batch(with_features=True)
    assert L.lib().lgn_sync_pipe_status(r.handle, 0) == L._lib.E_CAPACITY
    r.close()
    r = L.Runner(d.n_nodes, d.dim, B, fanout, rng_mode=L.RNG_PHILOX, rng_seed=9)
    r.bind_topology(ip, ix); r.bind_features(ft)
    cap = r.capacity
    mine = dict(ids=L.DevArray.zeros((cap,), np.int32), features=L.DevArray.zeros((cap, d.dim), np.float32), labels=L.DevArray.zeros((B,), np.int32),
                agg_src=L.DevArray.zeros((cap,), np.int32), agg_dst=L.DevArray.zeros((cap,), np.int32),
                node_counter=L.DevArray.zeros((16,), np.int32), edge_counter=L.DevArray.zeros((16,), np.int32))
    v = L._lib.BatchView()
    for k, a in mine.items():
        setattr(v, k, a.ptr)
    v.capacity, v.max_rows = cap, cap
    L._lib.check(L.lib().lgn_attach_buffers(r.handle, 0, C.byref(v)), "lgn_attach_buffers")
    r.batch_from_host(seeds, d.labels[seeds], step=0, pipe=0)
    r.run_batch(with_features=True)
    assert L.lib().lgn_sync_pipe_status(r.handle, 0) == 0
    want = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=9).sample(seeds, step=0)
    total, n_e = int(want["nc"][0]), int(want["ec"][0])
    assert np.array_equal(mine["node_counter"].numpy(), want["nc"]) and np.array_equal(mine["edge_counter"].numpy(), want["ec"])
    assert np.array_equal(mine["ids"].numpy(total), want["sampled_ids"][:total])
    assert np.array_equal(mine["agg_src"].numpy(n_e), want["agg_src_off"][:n_e]) and np.array_equal(mine["agg_dst"].numpy(n_e), want["agg_dst_off"][:n_e])
    assert np.array_equal(mine["features"].numpy(total).view(np.uint32), d.features[want["sampled_ids"][:total]].view(np.uint32))
    assert np.array_equal(mine["labels"].numpy(), d.labels[seeds])
    bad = L._lib.BatchView()
    bad.ids, bad.capacity = mine["ids"].ptr, B - 1                       # cannot hold the seeds
    assert L.lib().lgn_attach_buffers(r.handle, 1, C.byref(bad)) == L._lib.E_CAPACITY
    bad = L._lib.BatchView()
    bad.agg_src, bad.capacity = mine["agg_src"].ptr + 4, cap             # 128-bit accesses need 16-byte alignment
    assert L.lib().lgn_attach_buffers(r.handle, 1, C.byref(bad)) == -1
    r.close()
