"""Trainer-side checker shared by the IPC tests: consumes every batch the server publishes for one
device through the drop-in ipc_service module and compares it with the CPU oracle.  Runnable as a
script (one process per GPU, like the reference trainers): python _ipc_check.py <dev> <parts> ..."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def consume_and_check(dev, parts, n_nodes, avg_deg, dim, n_class, B, epochs, fanout, rng, seed):
    import torch
    import legion_b200 as L
    from legion_b200 import ipc_service
    from oracle import oracle as O
    d = L.synth.make_dataset(n_nodes, avg_deg, dim, n_class=n_class)
    torch.cuda.set_device(dev)
    ipc_service.initialize()
    steps = ipc_service.get_steps()
    split = lambda ids: [ids[(ids % parts) == p].astype(np.int32) for p in range(parts)]      # GPUGraphStore.cu:332-376
    tr, va, te = split(d.train_ids), split(d.valid_ids), split(d.test_ids)
    st = O.coordinate([len(x) for x in tr], [len(x) for x in va], [len(x) for x in te], B, epochs)
    assert steps == [st.train_step, st.valid_step, st.test_step], (steps, st.train_step, st.valid_step, st.test_step)
    lists = [(tr[dev], d.labels[tr[dev]]), (va[dev], d.labels[va[dev]]), (te[dev], d.labels[te[dev]])]
    mode_batch = [B, st.valid_batch[dev], st.test_batch[dev]]
    smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_MINSTD if rng == "minstd" else O.RNG_PHILOX, rng_seed=seed)
    for g in range(st.max_step):
        mode, local = O.mode_of_step(st, epochs, g), O.local_batch_id(st, epochs, g)
        seeds, slab = O.batch_generate(lists[mode][0], lists[mode][1].astype(np.int32), mode_batch[mode], local)
        # Philox position used by the server (server.cpp RunOnce): epoch of the global batch id (+1: epoch 0 is the
        # presampling pass), step = batch id inside the epoch + a per-mode offset
        per_epoch = st.train_step + st.valid_step
        is_test = g >= per_epoch * epochs
        epoch = (epochs if is_test else (g // per_epoch if per_epoch else 0)) + 1
        off = 0 if mode == 0 else (st.train_step if mode == 1 else per_epoch)
        want = smp.sample(seeds, step=local + off, epoch=epoch)
        nc, ec = want["nc"], want["ec"]
        if len(fanout) != 2:        # additive k-hop consumer API (the reference's get_next is hard-wired to two hops)
            H = len(fanout)
            ids, feats, labels, blocks = ipc_service.get_next_k(d.dim, H)
            sizes = ipc_service.get_block_sizes_k(H)
            total = int(nc[7 + 2 * (H - 1)])
            assert np.array_equal(ids.cpu().numpy(), want["sampled_ids"][:total]) and np.array_equal(labels.cpu().numpy(), slab)
            assert np.array_equal(feats.cpu().numpy().view(np.uint32), d.features[want["sampled_ids"][:total]].view(np.uint32))
            for layer, ((bs, bd), (n_src, n_dst)) in enumerate(zip(blocks, sizes)):
                h = H - 1 - layer
                e = int(ec[3 + h])
                assert (n_src, n_dst) == (nc[7 + 2 * h], nc[5 + 2 * h])
                assert np.array_equal(bs.cpu().numpy(), want["agg_src_off"][:e]) and np.array_equal(bd.cpu().numpy(), want["agg_dst_off"][:e])
                assert e == 0 or (int(bs.max()) < n_src and int(bd.max()) < n_dst)      # dst nodes are a prefix of src nodes
            ipc_service.synchronize()
            continue
        ids, feats, labels, b1s, b1d, b2s, b2d = ipc_service.get_next(d.dim)
        sizes = ipc_service.get_block_size()
        assert sizes == [nc[9], nc[7], nc[7], nc[5]], (g, sizes, nc)
        assert np.array_equal(ids.cpu().numpy(), want["sampled_ids"][:nc[9]]), g
        assert np.array_equal(labels.cpu().numpy(), slab), g
        assert np.array_equal(b1s.cpu().numpy(), want["agg_src_off"][:ec[4]]) and np.array_equal(b1d.cpu().numpy(), want["agg_dst_off"][:ec[4]])
        assert np.array_equal(b2s.cpu().numpy(), want["agg_src_off"][:ec[3]]) and np.array_equal(b2d.cpu().numpy(), want["agg_dst_off"][:ec[3]])
        assert np.array_equal(feats.cpu().numpy().view(np.uint32), d.features[want["sampled_ids"][:nc[9]]].view(np.uint32)), g
        assert feats.is_cuda and feats.data_ptr() != 0                          # zero-copy view of server memory
        ipc_service.synchronize()
    ipc_service.finalize()
    return st.max_step


if __name__ == "__main__":
    a = json.loads(sys.argv[1])
    n = consume_and_check(**a)
    print(f"device {a['dev']}: {n} batches bit-exact")
