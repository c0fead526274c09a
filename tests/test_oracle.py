"""CPU suite: pins the oracle (golden vectors generated from the CUDA toolkit's own thrust /
libcu++ headers by oracle/pins/*.cpp), cross-checks it against an independent pure-Python
restatement of the reference kernels on small cases, and checks the host logic of the product
library that needs no GPU (symbol table, step arithmetic)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_minstd_matches_thrust_fixture():
    fx = json.load(open(os.path.join(GOLDEN, "minstd_pin.json")))
    assert fx["first_raw"] == 48271
    for deg, row in zip(fx["degs"], fx["picks"]):
        got = [O.minstd_pick(i, deg) for i in fx["idx"]]
        assert got == row, f"deg {deg}"


def test_minstd_known_answer_from_survey():
    # deg=15, idx 0..5 -> 0,1,9,13,14,2
    assert [O.minstd_pick(i, 15) for i in range(6)] == [0, 1, 9, 13, 14, 2]


def test_philox_matches_libcudacxx_fixture_and_random123_kat():
    fx = json.load(open(os.path.join(GOLDEN, "philox_pin.json")))
    for c in fx["cases"]:
        assert O.philox4x32_10(c["ctr"], c["key"]) == c["out"]
    # Random123 kat_vectors: philox4x32 10 rounds
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # C++26 [rand.predef]: 10000th invocation of a default philox4x32 is 1955073260
    x = None
    for blk in range(2500):
        x = O.philox4x32_10([blk, 0, 0, 0], [20111115, 0])
    assert fx["default_10000th"] == 1955073260 and x[3] == 1955073260


# ---- independent pure-Python restatement of the reference kernels (small cases only) ----
def _py_pick(mode, idx, hop, step, seed, deg, f, k, epoch=0):
    if mode == O.RNG_MINSTD:
        x = pow(48271, idx + 1, 2147483647)                       # minstd_rand().discard(idx) then one draw
        return int((x - 1) / 2147483646.0 * deg)                  # uniform_int_distribution via double
    if deg <= f:
        return k
    return (O.philox4x32_10([idx & 0xffffffff, epoch, hop, step], [seed & 0xffffffff, seed >> 32])[0] * deg) >> 32   # counter = (slot, epoch, hop, step)


def _py_reference_batch(indptr, indices, seeds, fanout, mode, seed, step, epoch=0):
    """Kernels.cu:68-96 (batch_generator), 342-448 (sampler), 450-463 (construct_graph),
    112-150 (update_counter) executed slot by slot in idx order."""
    ids = [int(s) for s in seeds]
    pos = {}
    for i, s in enumerate(ids):
        if s >= 0 and s not in pos:
            pos[s] = i
    B = len(ids)
    nc, ec = [0] * 16, [0] * 16
    nc[0] = nc[2] = nc[4] = B
    src_ids, dst_ids = [], []
    frontier = list(ids)
    for h, f in enumerate(fanout):
        new_e_src, new_e_dst, n_new = [], [], 0
        for idx in range(len(frontier) * f):
            s, k = frontier[idx // f], idx % f
            if s < 0:
                continue
            start, deg = int(indptr[s]), int(indptr[s + 1] - indptr[s])
            if k >= deg:
                continue
            d = int(indices[start + _py_pick(mode, idx, h, step, seed, deg, f, k, epoch)])
            if d < 0:
                continue
            if d not in pos:
                pos[d] = len(ids)
                ids.append(d)
                n_new += 1
            new_e_src.append(d)
            new_e_dst.append(s)
        nc[5 + 2 * h] = nc[3 + 2 * h] + nc[4 + 2 * h]
        nc[6 + 2 * h] = n_new
        nc[7 + 2 * h] = nc[5 + 2 * h] + n_new
        nc[0] += n_new
        nc[2] = len(new_e_src)
        ec[2] = ec[0]
        ec[0] += len(new_e_src)
        ec[3 + h] = ec[0]
        src_ids += new_e_src
        dst_ids += new_e_dst
        frontier = new_e_src
    return dict(nc=np.array(nc, np.int32), ec=np.array(ec, np.int32), sampled_ids=np.array(ids, np.int32),
                agg_src_ids=np.array(src_ids, np.int32), agg_dst_ids=np.array(dst_ids, np.int32),
                agg_src_off=np.array([pos[x] for x in src_ids], np.int32),
                agg_dst_off=np.array([pos[x] for x in dst_ids], np.int32))


@pytest.mark.parametrize("mode", [O.RNG_MINSTD, O.RNG_PHILOX])
@pytest.mark.parametrize("fanout", [[5, 3], [4, 3, 2], [30]])
def test_oracle_equals_python_restatement(small, mode, fanout):
    d = small
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=mode, rng_seed=99)
    for step, seeds in enumerate([d.train_ids[:48], np.array([3, -1, 3, 17, -1], np.int32), np.array([], np.int32)]):
        got = smp.sample(seeds, step=step)
        want = _py_reference_batch(d.indptr, d.indices, seeds, fanout, mode, 99, step)
        total, n_e = int(want["nc"][0]), int(want["ec"][0])
        assert np.array_equal(got["nc"], want["nc"]) and np.array_equal(got["ec"], want["ec"])
        assert np.array_equal(got["sampled_ids"][:total], want["sampled_ids"])
        for k in ("agg_src_ids", "agg_dst_ids", "agg_src_off", "agg_dst_off"):
            assert np.array_equal(got[k][:n_e], want[k]), k
    assert (smp.pos == -1).all()


def test_counter_semantics_two_hops(c1):
    """Appendix A of SURVEY.md / update_counter (Kernels.cu:112-150) and the trainer's view
    (ipc_service.cpp:60-72): dst nodes are a prefix of src nodes in both blocks."""
    d = c1
    smp = O.Sampler(d.indptr, d.indices, [25, 10], rng_mode=O.RNG_MINSTD)
    out = smp.sample(d.train_ids[:1024])
    nc, ec = out["nc"], out["ec"]
    B = 1024
    n1, n2, e1, e2 = nc[6], nc[8], ec[3], ec[4] - ec[3]
    assert list(nc[:10]) == [B + n1 + n2, 0, e2, 0, B, B, n1, B + n1, n2, B + n1 + n2]
    assert list(ec[:5]) == [e1 + e2, 0, e1, e1, e1 + e2]
    ids = out["sampled_ids"][:nc[9]]
    assert len(np.unique(ids)) == len(ids)
    so, do = out["agg_src_off"][:ec[4]], out["agg_dst_off"][:ec[4]]
    assert np.array_equal(ids[so], out["agg_src_ids"][:ec[4]]) and np.array_equal(ids[do], out["agg_dst_ids"][:ec[4]])
    assert do[:e1].max() < B and so[:e1].max() < nc[7]            # block2: #src nc7, #dst nc5
    assert do.max() < nc[7] and so.max() < nc[9]                  # block1: #src nc9, #dst nc7
    # hop-2 frontier is hop-1's edge list incl. duplicates (Kernels.cu:371-373)
    assert set(out["agg_dst_ids"][e1:ec[4]]) <= set(out["agg_src_ids"][:e1])


def test_threaded_oracle_is_identical(c1):
    d = c1
    for mode in (O.RNG_MINSTD, O.RNG_PHILOX):
        a = O.Sampler(d.indptr, d.indices, [25, 10], rng_mode=mode, rng_seed=5)
        b = O.Sampler(d.indptr, d.indices, [25, 10], rng_mode=mode, rng_seed=5, n_threads=4)
        x, y = a.sample(d.train_ids[:1024], step=3), b.sample(d.train_ids[:1024], step=3)
        assert all(np.array_equal(x[k], y[k]) for k in x)


def test_philox_epoch_word(small):
    """counter word 1 = epoch: another epoch redraws other neighbourhoods of the same seeds; epoch 0 is the
    round-1 stream (slot index in word 0, zero in word 1)."""
    d = small
    smp = O.Sampler(d.indptr, d.indices, [3, 2], rng_mode=O.RNG_PHILOX, rng_seed=5)
    seeds = d.train_ids[:40]
    outs = []
    for epoch in (0, 1, 7):
        got = smp.sample(seeds, step=2, epoch=epoch)
        want = _py_reference_batch(d.indptr, d.indices, seeds, [3, 2], O.RNG_PHILOX, 5, 2, epoch=epoch)
        n_e = int(want["ec"][0])
        assert np.array_equal(got["ec"], want["ec"]) and np.array_equal(got["agg_src_ids"][:n_e], want["agg_src_ids"])
        outs.append(got["agg_src_ids"][:n_e].copy())
    assert not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[1], outs[2])
    assert O.philox_pick(123, 1, 9, 5, 1000, epoch=0) == (O.philox4x32_10([123, 0, 1, 9], [5, 0])[0] * 1000) >> 32
    assert O.philox_pick(123, 1, 9, 5, 1000, epoch=4) == (O.philox4x32_10([123, 4, 1, 9], [5, 0])[0] * 1000) >> 32


def test_philox_full_neighbourhood_and_step_dependence(small):
    import legion_b200 as L
    d = L.synth.make_dataset(5_000, 8.0, 4, kmax=3, with_features=False)
    deg = np.diff(d.indptr)
    f = int(deg.max())
    smp = O.Sampler(d.indptr, d.indices, [f], rng_mode=O.RNG_PHILOX)
    out = smp.sample(np.arange(100, dtype=np.int32))
    want = np.concatenate([d.indices[d.indptr[s]:d.indptr[s + 1]] for s in range(100)])
    assert np.array_equal(out["agg_src_ids"][:out["ec"][0]], want)
    d = small
    smp2 = O.Sampler(d.indptr, d.indices, [2, 2], rng_mode=O.RNG_PHILOX, rng_seed=1)
    a, b = smp2.sample(d.train_ids[:64], step=0), smp2.sample(d.train_ids[:64], step=1)
    assert not np.array_equal(a["agg_src_ids"], b["agg_src_ids"])  # the reference redraws identically every batch
    m = O.Sampler(d.indptr, d.indices, [2, 2], rng_mode=O.RNG_MINSTD)
    a, b = m.sample(d.train_ids[:64], step=0), m.sample(d.train_ids[:64], step=1)
    assert np.array_equal(a["agg_src_ids"], b["agg_src_ids"])      # Appendix B.1


def test_hotness_order_placement_and_gather(small):
    d = small
    smp = O.Sampler(d.indptr, d.indices, [5, 3], rng_mode=O.RNG_PHILOX)
    smp.enable_hotness()
    seen = np.zeros(d.n_nodes, np.int64)
    edges = 0
    for step in range(4):
        o = smp.sample(d.train_ids[step * 64:(step + 1) * 64], step=step)
        np.add.at(seen, o["sampled_ids"][:o["nc"][0]], 1)
        edges += int(o["ec"][0])
    assert np.array_equal(seen, smp.node_hotness) and smp.topo_hotness.sum() == edges
    order = O.hot_order(smp.node_hotness)
    key = smp.node_hotness[order].astype(np.int64)
    assert np.all(np.diff(key) <= 0)
    ties = np.flatnonzero(np.diff(key) == 0)
    assert np.all(order[ties] < order[ties + 1])                    # (count desc, id asc)
    kg, cap = 4, 500
    slot = O.place(order, cap, kg)
    for i in (0, 1, 5, 1999):
        assert slot[order[i]] == (i % kg) * cap + i // kg           # GPUCache.cu:106
    assert (slot[order[cap * kg:]] == -1).all()
    shards = [O.fill_feature_shard(order, cap, kg, j, d.features) for j in range(kg)]
    assert np.array_equal(shards[1][7], d.features[order[7 * kg + 1]])   # GPUCache.cu:202
    ids = np.concatenate([order[:50], order[-50:], [-1]]).astype(np.int32)
    out = np.zeros((len(ids), d.dim), np.float32)
    tiers = O.gather(ids, 0, len(ids), slot, cap, shards, d.features, out, tiers=True)
    assert np.array_equal(out[:100], d.features[ids[:100]]) and (out[100] == 0).all()
    assert tiers.sum() == 100 and tiers[kg] == 50
    ip, ix = O.fill_topo_shard(order, cap, kg, 2, d.indptr, d.indices)
    t = 11
    node = order[t * kg + 2]
    assert np.array_equal(ix[ip[t]:ip[t + 1]], d.indices[d.indptr[node]:d.indptr[node + 1]])


@pytest.mark.parametrize("n,n_repl,n_part,kg", [(1000, 100, 400, 4), (96 * 7, 96 * 7, 0, 1), (5000, 0, 5000, 8), (777, 10, 20, 2), (97, 0, 0, 3)])
def test_compact_placement_records_decode_to_the_slot_table(n, n_repl, n_part, kg):
    """the record format of include/legion_b200.h (lgn_place_compact) decoded word by word, the way the gather kernel does
    it (prefix + popcount), must give the slot table of the same placement; the cached SET is the hottest ranks."""
    rng = np.random.default_rng(n)
    order = O.hot_order(rng.integers(0, 9, n).astype(np.uint32))
    cap = n_repl + (n_part + kg - 1) // kg + 1
    me = kg - 1
    words, slot = O.place_compact(order, n_repl, n_part, kg, me, cap)
    w = words.reshape(-1, 8)
    assert w.shape[0] == (n + 95) // 96
    pc = lambda x: bin(int(x)).count("1")
    for nid in range(n):
        r, j = divmod(nid, 96)
        wi, bit = j >> 5, j & 31
        below = (1 << bit) - 1
        want = -1
        if (int(w[r, 2 + wi]) >> bit) & 1:
            want = me * cap + int(w[r, 0]) + sum(pc(w[r, 2 + k]) for k in range(wi)) + pc(int(w[r, 2 + wi]) & below)
        elif (int(w[r, 5 + wi]) >> bit) & 1:
            q = int(w[r, 1]) + sum(pc(w[r, 5 + k]) for k in range(wi)) + pc(int(w[r, 5 + wi]) & below)
            want = (q % kg) * cap + n_repl + q // kg
        assert slot[nid] == want
    cached = np.flatnonzero(slot >= 0)
    assert set(cached.tolist()) == set(order[:n_repl + n_part].tolist())
    # every (GPU, row) is used at most once, replicated rows are the same on every GPU
    parts = slot[slot >= 0]
    assert len(np.unique(parts)) == len(parts)
    _, slot0 = O.place_compact(order, n_repl, n_part, kg, 0, cap)
    assert np.array_equal(slot0[order[:n_repl]] % cap, slot[order[:n_repl]] % cap)


def test_batch_generate_quirk():
    ids = np.arange(100, 137, dtype=np.int32)
    lab = ids % 5
    a, _ = O.batch_generate(ids, lab, 16, 0)
    assert np.array_equal(a, ids[:16])
    b, lb = O.batch_generate(ids, lab, 16, 2)         # clamped to 5 seeds; stride = clamped size (Kernels.cu:224-227)
    assert len(b) == 5 and np.array_equal(b, ids[10:15]) and np.array_equal(lb, lab[10:15])


def test_step_arithmetic_matches_product_library():
    """IPCEnv::Coordinate / GetCurrentMode / GetLocalBatchId (CUDA_IPC_Service.cu:66-134, 219-259):
    oracle vs the C-ABI's host-only functions (no GPU needed)."""
    import legion_b200 as L
    for parts, B, ep in [(1, 8000, 10), (8, 8000, 3), (4, 100, 2)]:
        rng = np.random.default_rng(parts)
        tr = rng.integers(5 * B, 9 * B, parts).tolist()
        va = rng.integers(600, 5000, parts).tolist()
        te = rng.integers(600, 5000, parts).tolist()
        so = O.coordinate(tr, va, te, B, ep)
        sl = L.coordinate(tr, va, te, B, ep)
        assert (so.train_step, so.valid_step, so.test_step, so.max_step) == (sl.train_step, sl.valid_step, sl.test_step, sl.max_step)
        assert so.train_step == (min(tr) - 1) // B and so.valid_step == (max(va) - 1) // 512 + 1
        assert list(so.valid_batch)[:parts] == list(sl.valid_batch)[:parts] == [(v - 1) // so.valid_step + 1 for v in va]
        import ctypes as C
        for g in range(so.max_step):
            assert O.mode_of_step(so, ep, g) == L.lib().lgn_mode_of_step(C.byref(sl), ep, g)
            assert O.local_batch_id(so, ep, g) == L.lib().lgn_local_batch_id(C.byref(sl), ep, g)


def test_library_exports_every_declared_symbol():
    import legion_b200 as L
    lib = L.lib()
    names = L._lib.declared_symbols()
    assert len(names) > 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/legion_b200.h but not exported"


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(__file__))
    pkg = os.path.join(root, "legion-1_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_synth_numpy_equals_torch():
    import torch
    import legion_b200 as L
    a = L.synth.make_dataset(30_000, 12.0, 32)
    b = L.synth.make_dataset(30_000, 12.0, 32, backend="torch", device="cpu", dmin_fp=a.dmin_fp)
    assert np.array_equal(a.indptr, b.indptr.numpy()) and np.array_equal(a.indices, b.indices.numpy())
    assert np.array_equal(a.features.view(np.uint32), b.features.numpy().view(np.uint32))
    assert np.array_equal(a.train_ids, b.train_ids.numpy())
    assert abs(a.n_edges / a.n_nodes - 12.0) < 0.5
    assert torch.equal(b.labels, torch.from_numpy(a.labels))


def test_dataset_files_round_trip(tmp_path):
    """the on-disk format the reference loader reads (GPUGraphStore.cu:254-301): raw little-endian arrays with the
    reference's file names, sizes implied by the 11-field meta_config; write -> read must be bit-exact."""
    import legion_b200 as L
    from legion_b200 import dataset_io
    d = L.synth.make_dataset(3000, 6.0, 12, n_class=5)
    data = str(tmp_path / "data")
    dataset_io.write_dataset(data, d)
    line = dataset_io.write_meta_config(str(tmp_path), data, d, 64, 10**6, 2)
    f = line.split()
    assert len(f) == 11 and f[0].endswith("/") and [int(x) for x in f[1:8]] == [64, d.n_nodes, d.n_edges, d.dim, len(d.train_ids), len(d.valid_ids), len(d.test_ids)]
    sizes = {"edge_src": 8 * (d.n_nodes + 1), "edge_dst": 4 * d.n_edges, "features": 4 * d.n_nodes * d.dim, "labels": 4 * d.n_nodes,
             "trainingset": 4 * len(d.train_ids), "validationset": 4 * len(d.valid_ids), "testingset": 4 * len(d.test_ids)}
    for name, n in sizes.items():
        assert os.path.getsize(os.path.join(data, name)) == n, name
    back = dataset_io.read_dataset(data, d.n_nodes, d.n_edges, d.dim, len(d.train_ids), len(d.valid_ids), len(d.test_ids))
    for k in ("indptr", "indices", "labels", "train_ids", "valid_ids", "test_ids"):
        assert np.array_equal(getattr(back, k), getattr(d, k)), k
    assert np.array_equal(back.features.view(np.uint32), d.features.view(np.uint32))


def test_launcher_writes_reference_meta_config(tmp_path, monkeypatch, capsys):
    """legion_server.py keeps the reference launcher's command line (legion_server.py:72-85): dataset table, 11-field
    meta_config (path batch vertices edges dim train valid test cache epochs partition) and the clique-mode rule."""
    import legion_b200  # noqa: F401
    from legion_b200 import legion_server
    monkeypatch.chdir(tmp_path)
    assert legion_server.main(["--dataset", "PR", "--dataset_path", "/data", "--gpu_number", "4", "--epoch", "3", "--dry_run"]) == 0
    f = open(tmp_path / "meta_config").read().split()
    assert f == ["/data/products/", "8000", "2449029", "123718280", "100", "196615", "39323", "2213091", "38000000000", "3", "0"]
    assert capsys.readouterr().out.split()[-2:] == ["4", "2"]          # NVSwitch: one clique over all 4 GPUs (the reference pairs them: mode 1)
    assert legion_server.main(["--dataset", "PA", "--gpu_number", "8", "--cache_agg_mode", "3", "--usenvlink", "0", "--dry_run"]) == 0
    f = open(tmp_path / "meta_config").read().split()
    assert f[2:5] == ["111059956", "1615685872", "128"] and f[-1] == "1"
    assert capsys.readouterr().out.split()[-2:] == ["8", "3"]
    assert legion_server.main(["--dataset", "nope", "--dry_run"]) == 2
