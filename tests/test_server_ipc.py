"""End to end through every boundary layer: the `legion` server binary (reference-shaped Server /
Runner / Operator classes over the C-ABI, its dataset loader and meta_config parser) in one process,
the drop-in `ipc_service` module in another, talking through the reference's wire format (shm +
named semaphores + CUDA IPC handles).  Every batch of every mode is compared with the CPU oracle."""
import os
import subprocess
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEGION = os.path.join(ROOT, "legion-1_b200", "_build", "legion")


def _start_server(workdir, ngpu, agg_mode, env):
    log = open(os.path.join(workdir, "server.log"), "w")
    p = subprocess.Popen([LEGION, str(ngpu), str(agg_mode)], cwd=workdir, stdout=log, stderr=subprocess.STDOUT,
                         env={**os.environ, **env})
    t0 = time.time()
    while time.time() - t0 < 120:
        if p.poll() is not None:
            raise RuntimeError("server exited early:\n" + open(os.path.join(workdir, "server.log")).read()[-3000:])
        if "System is ready for serving" in open(os.path.join(workdir, "server.log")).read():
            return p
        time.sleep(0.2)
    p.kill()
    raise RuntimeError("server did not become ready:\n" + open(os.path.join(workdir, "server.log")).read()[-3000:])


@pytest.mark.parametrize("rng,cache_mem,runner", [("minstd", 800_000, "fast"), ("philox", 10**9, "fast"), ("philox", 800_000, "ops")])
def test_server_to_trainer_roundtrip(tmp_path, rng, cache_mem, runner):
    import torch
    import legion_b200 as L
    from legion_b200 import dataset_io, ipc_service
    from oracle import oracle as O
    if not os.path.exists(LEGION):
        pytest.skip("legion binary not built")
    d = L.synth.make_dataset(12_000, 10.0, 24, n_class=5)
    B, epochs, fanout = 200, 2, [25, 10]
    data_dir, work = str(tmp_path / "data"), str(tmp_path)
    dataset_io.write_dataset(data_dir, d)
    dataset_io.write_meta_config(work, data_dir, d, B, cache_mem, epochs)
    srv = _start_server(work, 1, 0, {"LEGION_RNG": rng, "LEGION_SEED": "77", "LEGION_RUNNER": runner})
    try:
        torch.cuda.set_device(0)
        ipc_service.initialize()
        steps = ipc_service.get_steps()
        lists = [(d.train_ids, d.labels[d.train_ids]), (d.valid_ids, d.labels[d.valid_ids]), (d.test_ids, d.labels[d.test_ids])]
        st = O.coordinate([len(d.train_ids)], [len(d.valid_ids)], [len(d.test_ids)], B, epochs)
        assert steps == [st.train_step, st.valid_step, st.test_step]
        mode_batch = [B, st.valid_batch[0], st.test_batch[0]]
        mode_o = O.RNG_MINSTD if rng == "minstd" else O.RNG_PHILOX
        smp = O.Sampler(d.indptr, d.indices, fanout, rng_mode=mode_o, rng_seed=77)
        for g in range(st.max_step):
            mode, local = O.mode_of_step(st, epochs, g), O.local_batch_id(st, epochs, g)
            ids, feats, labels, b1s, b1d, b2s, b2d = ipc_service.get_next(d.dim)
            sizes = ipc_service.get_block_size()
            seeds, slab = O.batch_generate(lists[mode][0].astype(np.int32), lists[mode][1].astype(np.int32), mode_batch[mode], local)
            want = smp.sample(seeds, step=local)
            nc, ec = want["nc"], want["ec"]
            assert sizes == [nc[9], nc[7], nc[7], nc[5]], (g, sizes, nc)
            assert np.array_equal(ids.cpu().numpy(), want["sampled_ids"][:nc[9]]), g
            assert np.array_equal(labels.cpu().numpy(), slab), g
            assert np.array_equal(b1s.cpu().numpy(), want["agg_src_off"][:ec[4]]) and np.array_equal(b1d.cpu().numpy(), want["agg_dst_off"][:ec[4]])
            assert np.array_equal(b2s.cpu().numpy(), want["agg_src_off"][:ec[3]]) and np.array_equal(b2d.cpu().numpy(), want["agg_dst_off"][:ec[3]])
            assert np.array_equal(feats.cpu().numpy().view(np.uint32), d.features[want["sampled_ids"][:nc[9]]].view(np.uint32)), g
            assert feats.data_ptr() != 0 and feats.is_cuda                      # zero-copy view of server memory
            ipc_service.synchronize()
        ipc_service.finalize()
        assert srv.wait(timeout=60) == 0
        log = open(os.path.join(work, "server.log")).read()
        assert "Server Stopped" in log
        if cache_mem < 10**9:
            assert "whole graph fits" not in log          # the restricted budget went through the cost model
    finally:
        if srv.poll() is None:
            srv.kill()
