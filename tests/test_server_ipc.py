"""End to end through every boundary layer: the `legion` server binary (reference-shaped Server /
Runner / Operator classes over the C-ABI, its dataset loader and meta_config parser) in one process,
the drop-in `ipc_service` module in another, talking through the reference's wire format (shm +
named semaphores + CUDA IPC handles).  Every batch of every mode is compared with the CPU oracle."""
import os
import subprocess
import sys
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


@pytest.mark.parametrize("rng,cache_mem,runner,fanout", [("minstd", 800_000, "fast", [25, 10]), ("philox", 10**9, "fast", [25, 10]),
                                                         ("philox", 800_000, "ops", [25, 10]), ("philox", 800_000, "fast", [15, 10, 5])])
def test_server_to_trainer_roundtrip(tmp_path, rng, cache_mem, runner, fanout):
    import legion_b200 as L
    from legion_b200 import dataset_io
    sys.path.insert(0, os.path.dirname(__file__))
    from _ipc_check import consume_and_check
    if not os.path.exists(LEGION):
        pytest.skip("legion binary not built")
    cfg = dict(n_nodes=12_000, avg_deg=10.0, dim=24, n_class=5)
    d = L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"])
    B, epochs = 200, 2
    data_dir, work = str(tmp_path / "data"), str(tmp_path)
    dataset_io.write_dataset(data_dir, d)
    dataset_io.write_meta_config(work, data_dir, d, B, cache_mem, epochs)
    srv = _start_server(work, 1, 0, {"LEGION_RNG": rng, "LEGION_SEED": "77", "LEGION_RUNNER": runner,
                                     "LEGION_FANOUT": ",".join(map(str, fanout))})
    try:
        n = consume_and_check(dev=0, parts=1, B=B, epochs=epochs, fanout=fanout, rng=rng, seed=77, **cfg)
        assert n > 10
        assert srv.wait(timeout=60) == 0
        log = open(os.path.join(work, "server.log")).read()
        assert "Server Stopped" in log
        if cache_mem < 10**9:
            assert "whole graph fits" not in log          # the restricted budget went through the cost model
    finally:
        if srv.poll() is None:
            srv.kill()


def test_trainer_example_trains_through_ipc(tmp_path):
    """the GraphSAGE example (reference trainer's structure, this repo's ipc_service + DGL stand-ins) consumes a
    whole run of the server: train / valid / test phases, loss finite, accuracy printed, both sides exit cleanly."""
    import legion_b200 as L
    from legion_b200 import dataset_io, legion_server
    cfg = dict(n_nodes=12_000, avg_deg=10.0, dim=24, n_class=5)
    d = L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"])
    data_dir, work = str(tmp_path / "data"), str(tmp_path)
    dataset_io.write_dataset(data_dir, d)
    cwd = os.getcwd()
    os.chdir(work)
    try:      # the launcher writes ./meta_config exactly like the reference's legion_server.py
        custom = ",".join(map(str, [data_dir, d.n_nodes, d.n_edges, d.dim, len(d.train_ids), len(d.valid_ids), len(d.test_ids)]))
        assert legion_server.main(["--custom", custom, "--train_batch_size", "200", "--epoch", "2", "--cache_memory", "1000000000",
                                   "--gpu_number", "1", "--dry_run"]) == 0
    finally:
        os.chdir(cwd)
    assert open(os.path.join(work, "meta_config")).read().split()[1:5] == ["200", str(d.n_nodes), str(d.n_edges), str(d.dim)]
    srv = _start_server(work, 1, 0, {"LEGION_RNG": "philox"})
    try:
        ex = os.path.join(ROOT, "examples", "train_graphsage.py")
        out = subprocess.run([sys.executable, ex, "--gpu", "0", "--features_num", str(d.dim), "--class_num", "5", "--hidden_dim", "32",
                              "--epoch", "2"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert out.stdout.count("Epoch:") == 2 and "Accuracy on test data" in out.stdout
        assert "nan" not in out.stdout.lower()
        assert srv.wait(timeout=60) == 0
    finally:
        if srv.poll() is None:
            srv.kill()


@pytest.mark.parametrize("cache_mem", [600_000, 10**9])
def test_two_gpu_clique_server(tmp_path, cache_mem):
    """`legion 2 1`: one NVLink clique of two GPUs -- seeds partitioned tid % 2, hotness summed across the
    clique, feature and topology shards interleaved over both GPUs and read through P2P loads."""
    import json
    import torch
    import legion_b200 as L
    from legion_b200 import dataset_io
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cfg = dict(n_nodes=12_000, avg_deg=10.0, dim=128 if cache_mem > 10**8 else 24, n_class=5)   # 128-d rows need the >48 KB smem opt-in on BOTH devices
    d = L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"])
    B, epochs = 100, 2
    data_dir, work = str(tmp_path / "data"), str(tmp_path)
    dataset_io.write_dataset(data_dir, d)
    dataset_io.write_meta_config(work, data_dir, d, B, cache_mem, epochs)
    srv = _start_server(work, 2, 1, {"LEGION_RNG": "philox", "LEGION_SEED": "5"})
    try:
        checker = os.path.join(os.path.dirname(__file__), "_ipc_check.py")
        procs = []
        for dev in range(2):      # one trainer process per GPU, like mp.spawn in legion_graphsage.py
            a = dict(dev=dev, parts=2, B=B, epochs=epochs, fanout=[25, 10], rng="philox", seed=5, **cfg)
            procs.append(subprocess.Popen([sys.executable, checker, json.dumps(a)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=300)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
            assert "bit-exact" in o
        assert srv.wait(timeout=60) == 0
    finally:
        if srv.poll() is None:
            srv.kill()
