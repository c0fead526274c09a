"""The reference's entry-point names (Kernels.cuh:24-93, GPU_Graph_Storage.cuh:38-39, GPU_Node_Storage.cuh:60-61, the
GPUCache / GPUMemoryPool / IPCEnv classes) exported by liblegion_b200.so: a runner written in the reference's style
against those names only (tests/abi/ref_style_runner.cpp) produces batches that equal the CPU oracle bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "legion-1_b200", "_build")
COMPAT = os.path.join(ROOT, "legion-1_b200", "csrc", "compat")
C_NAMES = ["d_alloc_space", "d_alloc_space_managed", "d_copy_2_h", "d_free_space", "SetGPUDevice", "GetGPUDevice", "host_alloc_space",
           "batch_generator_kernel", "GPU_Random_Sampling", "get_feature_kernel", "make_update_plan", "update_cache",
           "NewGPUMemoryGraphStorage", "NewGPUMemoryNodeStorage"]


def test_library_exports_the_reference_entry_points():
    lib = C.CDLL(os.path.join(BUILD, "liblegion_b200.so"))
    for name in C_NAMES:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(BUILD, "liblegion_b200.so")], capture_output=True, text=True).stdout
    for mangled in ("NewIPCEnv", "NewPreSCCacheController", "GPUCache18CandidateSelection", "GPUCache9CostModel", "GPUCache6FillUp",
                    "GPUMemoryPoolC"):
        assert mangled in out, mangled


@pytest.mark.gpu
def test_reference_style_runner_matches_the_oracle(tmp_path):
    import legion_b200 as L
    from legion_b200 import dataset_io
    from oracle import oracle as O
    exe = str(tmp_path / "ref_style_runner")
    cc = subprocess.run(["nvcc", "-O1", "-std=c++17", "-I", COMPAT, "-o", exe, os.path.join(ROOT, "tests", "abi", "ref_style_runner.cpp"),
                         "-L", BUILD, "-llegion_b200", "-Xlinker", "-rpath", "-Xlinker", BUILD], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    d = L.synth.make_dataset(9_000, 9.0, 20, n_class=7)
    data_dir = str(tmp_path / "data")
    dataset_io.write_dataset(data_dir, d)
    B, n_batches = 128, 5
    for cache_bytes in (10**9, 300_000):             # everything cached / most rows served from the host tier
        out_path = str(tmp_path / f"out_{cache_bytes}.bin")
        run = subprocess.run([exe, data_dir, str(d.n_nodes), str(d.n_edges), str(d.dim), str(B), str(n_batches), str(cache_bytes), out_path],
                             capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
        raw = np.fromfile(out_path, np.int32)
        assert raw[0] == n_batches
        pos = 4
        train = np.arange(0, d.n_nodes, 3, dtype=np.int32)
        smp = O.Sampler(d.indptr, d.indices, [25, 10], rng_mode=O.RNG_MINSTD)       # the reference's stream (LEGION_RNG unset)
        for it in range(n_batches):
            seeds = train[it * B:(it + 1) * B]
            want = smp.sample(seeds, step=it)
            nc, ec = raw[pos:pos + 16], raw[pos + 16:pos + 32]
            pos += 32
            assert np.array_equal(nc, want["nc"]) and np.array_equal(ec, want["ec"]), it
            total, n_e, nb = int(nc[9]), int(ec[4]), int(nc[4])
            ids = raw[pos:pos + total]; pos += total
            lab = raw[pos:pos + nb]; pos += nb
            so = raw[pos:pos + n_e]; pos += n_e
            do = raw[pos:pos + n_e]; pos += n_e
            ft = raw[pos:pos + total * d.dim].view(np.uint32).reshape(total, d.dim); pos += total * d.dim
            assert np.array_equal(ids, want["sampled_ids"][:total]) and np.array_equal(lab, d.labels[seeds])
            assert np.array_equal(so, want["agg_src_off"][:n_e]) and np.array_equal(do, want["agg_dst_off"][:n_e])
            assert np.array_equal(ft, d.features[ids].view(np.uint32))
        assert pos == len(raw)
        if cache_bytes < 10**9:
            assert 0 < raw[1] < d.n_nodes            # the cost model restricted the feature cache
