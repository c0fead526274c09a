"""world_size-2 gloo tests (CPU) of the one-process-per-GPU host logic: seed partitioning, the hotness
all-reduce, handle exchange order and the consistency of the shard placement across ranks."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import legion_b200 as L
    from legion_b200 import cluster
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        d = L.synth.make_dataset(6_000, 8.0, 16, n_class=3)
        B, fanout = 64, [5, 3]
        mine = cluster.partition_seeds(d.train_ids, world, rank)
        parts = [None] * world
        dist.all_gather_object(parts, mine.tolist())
        allp = np.concatenate([np.asarray(p) for p in parts])
        assert len(allp) == len(d.train_ids) and np.array_equal(np.sort(allp), np.sort(d.train_ids))      # disjoint cover
        assert all((np.asarray(p) % world == r).all() for r, p in enumerate(parts))
        steps = cluster.train_steps(len(mine), B, dist)
        assert steps == (min(len(p) for p in parts) - 1) // B                                             # CUDA_IPC_Service.cu:88
        # presampling on this rank's partition (oracle as the per-rank sampler stand-in)
        smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=O.RNG_PHILOX, rng_seed=9)
        smp.enable_hotness()
        for s in range(steps):
            smp.sample(mine[s * B:(s + 1) * B], step=s)
        local = smp.node_hotness.copy()
        t = torch.from_numpy(smp.node_hotness.view(np.int32))
        cluster.allreduce_hotness(dist, t)                                # the path's one collective
        hists = [None] * world
        dist.all_gather_object(hists, local.tolist())
        assert np.array_equal(smp.node_hotness, np.sum([np.asarray(h, np.uint32) for h in hists], axis=0, dtype=np.uint32))
        order = O.hot_order(smp.node_hotness)                              # every rank sorts the same histogram
        orders = [None] * world
        dist.all_gather_object(orders, order.tolist())
        assert all(o == orders[0] for o in orders)
        # placement: every hot rank lives on exactly one GPU, slot map and shard rows agree
        n_cached = 4_000
        cap = cluster.capacity_for(n_cached, world)
        slot = O.place(order, cap, world)
        my_ranks = cluster.shard_ranks(cap, world, rank, d.n_nodes)
        assert np.array_equal(slot[order[my_ranks]], rank * cap + np.arange(len(my_ranks)))
        assert all(cluster.slot_of_rank(int(i), cap, world) == slot[order[i]] for i in (0, 1, 2, 17, cap * world - 1))
        shard = O.fill_feature_shard(order, cap, world, rank, d.features)
        handles = cluster.exchange_handles(dist, bytes([rank]) * 64)       # stand-in for lgn_ipc_export's 64 bytes
        assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)
        # descriptor exchange (shareable handles of lgn_shared_alloc travel as file descriptors over Unix sockets):
        # every rank offers a temp file holding its rank, and must read every other rank's content through the duplicate
        import tempfile
        with tempfile.TemporaryFile() as tf:
            tf.write(b"shard-of-rank-%d" % rank)
            tf.flush()
            fds = cluster.exchange_fds(dist, tf.fileno(), tag="t")
            assert fds[rank] == tf.fileno() and len(fds) == world
            for j, fd in enumerate(fds):
                assert os.pread(fd, 64, 0) == b"shard-of-rank-%d" % j
                if j != rank:
                    os.close(fd)
        shards = [None] * world
        dist.all_gather_object(shards, shard)
        batch = smp.sample(mine[:B], step=0)
        total = int(batch["nc"][0])
        out = np.zeros((total, d.dim), np.float32)
        tiers = O.gather(batch["sampled_ids"], 0, total, slot, cap, shards, d.features, out, tiers=True)
        assert np.array_equal(out.view(np.uint32), d.features[batch["sampled_ids"][:total]].view(np.uint32))
        assert tiers[:world].sum() > 0 and tiers.sum() == total
        # compact placement (lgn_place_compact): the map is the same on every rank, replicated rows are local everywhere,
        # partitioned rows live on exactly one rank; every rank gathers the same bit-exact rows through its own slot view
        n_repl_c, cap_c = 1_000, 1_000 + cluster.capacity_for(2_500, world)
        n_repl_c, n_part_c = cluster.compact_split(d.n_nodes, n_repl_c, cap_c, world)
        assert (n_repl_c, n_part_c) == (1_000, 2_500 + (2_500 % world))
        words, slot_c = O.place_compact(order, n_repl_c, n_part_c, world, rank, cap_c)
        all_words = [None] * world
        dist.all_gather_object(all_words, words.tobytes())
        assert all(w == all_words[0] for w in all_words)
        shard_c = O.fill_feature_shard_compact(slot_c, cap_c, rank, d.features)
        shards_c = [None] * world
        dist.all_gather_object(shards_c, shard_c)
        out_c = np.zeros((total, d.dim), np.float32)
        tiers_c = O.gather(batch["sampled_ids"], 0, total, slot_c, cap_c, shards_c, d.features, out_c, tiers=True)
        assert np.array_equal(out_c.view(np.uint32), d.features[batch["sampled_ids"][:total]].view(np.uint32))
        hot = np.isin(batch["sampled_ids"][:total], order[:n_repl_c]).sum()
        assert tiers_c[rank] >= hot and tiers_c.sum() == total
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"
