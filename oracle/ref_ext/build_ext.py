"""Builds the REFERENCE's own trainer-side pieces, unmodified, from where they lie under /root/reference (test
infrastructure, like oracle/ref_harness):

  * pytorch_extension/{ipc_service.cpp, helper_multiprocess.cpp, ipc_cuda_kernel.cu} -> oracle/_ref/ext/ipc_service*.so
    (torch C++/CUDA extension, sm_100a), the module the three trainers import;
  * pytorch_extension/{legion_graphsage,legion_gcn,lp_sage}.py -> oracle/_ref/trainers/*.bin (byte-compiled, nothing
    edited), so that the unchanged trainers can be run on the GPU box, where /root/reference does not exist.

tests/test_reference_trainers.py points them at this repo's `legion` server.  Outputs only under oracle/_ref/
(git-ignored).  No reference source is copied into the repository."""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LEGION_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "_ref")


def build(verbose=False):
    ext_src = os.path.join(REF, "pytorch_extension")
    if not os.path.isdir(ext_src):
        return None
    os.makedirs(os.path.join(OUT, "trainers"), exist_ok=True)
    for name in ("legion_graphsage", "legion_gcn", "lp_sage"):      # ".bin": snapshot tools tend to drop *.pyc; CPython runs a
        # byte-compiled file of any extension (it looks at the magic number)
        py_compile.compile(os.path.join(ext_src, name + ".py"), cfile=os.path.join(OUT, "trainers", name + ".bin"), doraise=True)
    ext_dir = os.path.join(OUT, "ext")
    os.makedirs(ext_dir, exist_ok=True)
    so = [f for f in os.listdir(ext_dir) if f.startswith("ipc_service") and f.endswith(".so")]
    if so:
        return os.path.join(ext_dir, so[0])
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    # the image's default CXX (/opt/gcc/bin/g++) links libstdc++ statically; a second, private copy of iostreams inside
    # a Python extension crashes on the extension's first `std::cout << int` (ipc_cuda_kernel.cu:77).  Use the system compiler.
    if os.path.exists("/usr/bin/g++"):
        os.environ["CXX"], os.environ["CC"] = "/usr/bin/g++", "/usr/bin/gcc"
    from torch.utils.cpp_extension import load
    load(name="ipc_service", sources=[os.path.join(ext_src, f) for f in ("ipc_service.cpp", "helper_multiprocess.cpp", "ipc_cuda_kernel.cu")],
         extra_cflags=["-O2", "-std=c++17"], extra_cuda_cflags=["-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=ext_dir, verbose=verbose, is_python_module=False)
    so = [f for f in os.listdir(ext_dir) if f.startswith("ipc_service") and f.endswith(".so")]
    return os.path.join(ext_dir, so[0]) if so else None


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
