"""The reference's OWN trainers, unchanged, against this repo's server: `legion` (server.cpp over the C-ABI) in one
process, the reference's `ipc_service` torch extension built UNMODIFIED from /root/reference/pytorch_extension
(oracle/ref_ext/build_ext.py -> oracle/_ref/ext/) and its byte-compiled, unedited legion_graphsage.py /
legion_gcn.py (oracle/_ref/trainers/*.bin) in another, DGL / torchmetrics provided by the stand-ins under
legion-1_b200/shims (neither is installed here).  Proves the wire format (shm layout, semaphore protocol, CUDA IPC
handles, counter slots) against the consumer it was written for: ipc_cuda_kernel.cu:38-230, ipc_service.cpp:43-93,
legion_graphsage.py:72-89,149-168."""
import glob
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEGION = os.path.join(ROOT, "legion-1_b200", "_build", "legion")
EXT_DIR = os.path.join(ROOT, "oracle", "_ref", "ext")
TRAINERS = os.path.join(ROOT, "oracle", "_ref", "trainers")
SHIMS = os.path.join(ROOT, "legion-1_b200", "shims")


def _have_reference_build():
    return bool(glob.glob(os.path.join(EXT_DIR, "ipc_service*.so"))) and os.path.exists(os.path.join(TRAINERS, "legion_graphsage.bin"))


@pytest.mark.parametrize("trainer", ["legion_graphsage", "legion_gcn"])
def test_unchanged_reference_trainer_runs_against_the_server(tmp_path, trainer):
    import legion_b200 as L
    from legion_b200 import dataset_io
    sys.path.insert(0, os.path.dirname(__file__))
    from test_server_ipc import _start_server
    if not _have_reference_build():
        pytest.skip("reference extension / trainers not built (oracle/ref_ext/build_ext.py needs /root/reference)")
    if not os.path.exists(LEGION):
        pytest.skip("legion binary not built")
    cfg = dict(n_nodes=12_000, avg_deg=10.0, dim=24, n_class=5)
    d = L.synth.make_dataset(cfg["n_nodes"], cfg["avg_deg"], cfg["dim"], n_class=cfg["n_class"])
    B, epochs = 200, 2
    data_dir, work = str(tmp_path / "data"), str(tmp_path)
    dataset_io.write_dataset(data_dir, d)
    dataset_io.write_meta_config(work, data_dir, d, B, 10**9, epochs)
    srv = _start_server(work, 1, 0, {})        # the reference's defaults: minstd stream, fanout [25, 10]
    try:
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([EXT_DIR, SHIMS, os.environ.get("PYTHONPATH", "")]))
        env.pop("MASTER_ADDR", None); env.pop("MASTER_PORT", None)
        out = subprocess.run([sys.executable, os.path.join(TRAINERS, trainer + ".bin"), "--class_num", str(cfg["n_class"]),
                              "--features_num", str(cfg["dim"]), "--train_batch_size", str(B), "--hidden_dim", "32",
                              "--epoch", str(epochs), "--gpu_num", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=work)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        assert out.stdout.count("Epoch:") == epochs and "Accuracy on test data" in out.stdout, out.stdout[-2000:]
        assert "nan" not in out.stdout.lower()
        assert srv.wait(timeout=120) == 0
        assert "Server Stopped" in open(os.path.join(work, "server.log")).read()
    finally:
        if srv.poll() is None:
            srv.kill()
