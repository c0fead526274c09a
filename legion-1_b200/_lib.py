"""ctypes binding of the C-ABI (include/legion_b200.h).  The CUDA library is the only
compute path: if it is missing or a call fails this module raises, it never falls back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO_PATH = os.environ.get("LGN_LIBRARY") or os.path.join(_HERE, "_build", "liblegion_b200.so")      # LGN_LIBRARY: the debug build
HEADER = os.path.join(ROOT, "include", "legion_b200.h")

MAX_HOPS, MAX_PARTS, PIPELINE_DEPTH = 5, 8, 2
RNG_MINSTD, RNG_PHILOX = 0, 1
MODE_TRAIN, MODE_VALID, MODE_TEST = 0, 1, 2
E_CAPACITY = -4


class LegionError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("part", C.c_int32), ("n_nodes", C.c_int64), ("feat_dim", C.c_int32),
                ("batch_size", C.c_int32), ("n_hops", C.c_int32), ("fanout", C.c_int32 * MAX_HOPS),
                ("rng_mode", C.c_int32), ("rng_seed", C.c_uint64), ("max_feature_rows", C.c_int64),
                ("enable_hotness", C.c_int32), ("n_lanes", C.c_int32)]


class BatchView(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("features", C.c_void_p), ("labels", C.c_void_p), ("agg_src", C.c_void_p),
                ("agg_dst", C.c_void_p), ("node_counter", C.c_void_p), ("edge_counter", C.c_void_p),
                ("agg_src_ids", C.c_void_p), ("agg_dst_ids", C.c_void_p), ("capacity", C.c_int64),
                ("max_rows", C.c_int64)]


class Steps(C.Structure):
    _fields_ = [("train_step", C.c_int32), ("valid_step", C.c_int32), ("test_step", C.c_int32),
                ("max_step", C.c_int32), ("valid_batch", C.c_int32 * MAX_PARTS), ("test_batch", C.c_int32 * MAX_PARTS)]


def build(verbose=False):
    """compile liblegion_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8", "all", "debug"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise LegionError("building liblegion_b200.so failed")
    return SO_PATH


def declared_symbols():
    """every function name the public header declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lgn_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LegionError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the hot path)")
        L = C.CDLL(SO_PATH)
        L.lgn_error_string.restype = C.c_char_p
        L.lgn_last_cuda_error.restype = C.c_char_p
        L.lgn_capacity.restype = C.c_int64
        L.lgn_cmap_bytes.restype = C.c_int64
        L.lgn_max_ids.restype = C.c_int32
        L.lgn_mode_of_step.restype = C.c_int32
        L.lgn_local_batch_id.restype = C.c_int32
        L.lgn_launches_per_batch.restype = C.c_int32
        L.lgn_gather_kernel_name.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        L = lib()
        msg = L.lgn_error_string(rc).decode()
        if rc == -2:
            msg += ": " + L.lgn_last_cuda_error().decode()
        raise LegionError(f"{what} failed: {msg} (rc={rc})")


def _vp(x):
    if x is None:
        return None
    if isinstance(x, (DevArray, MappedHostArray)):
        return C.c_void_p(x.ptr)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):            # torch tensor
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    raise TypeError(type(x))


class DevArray:
    """a typed device allocation owned through lgn_device_alloc (cudaMalloc)."""

    def __init__(self, shape, dtype, ptr=None, owner=True):
        self.shape = tuple(np.atleast_1d(shape).tolist()) if not isinstance(shape, tuple) else shape
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self.owner = owner and ptr is None
        if ptr is None:
            p = C.c_void_p()
            check(lib().lgn_device_alloc(C.byref(p), C.c_int64(self.nbytes)), "lgn_device_alloc")
            ptr = p.value
        self.ptr = ptr

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype)
        if a.nbytes:
            check(lib().lgn_copy_h2d(C.c_void_p(d.ptr), a.ctypes.data_as(C.c_void_p), C.c_int64(a.nbytes)), "h2d")
        return d

    @classmethod
    def zeros(cls, shape, dtype):
        d = cls(shape, dtype)
        if d.nbytes:
            check(lib().lgn_memset_d(C.c_void_p(d.ptr), 0, C.c_int64(d.nbytes)), "memset")
        return d

    def numpy(self, count=None):
        shape = self.shape if count is None else (count,) + self.shape[1:]
        out = np.empty(shape, self.dtype)
        if out.nbytes:
            check(lib().lgn_copy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), C.c_int64(out.nbytes)), "d2h")
        return out

    def view(self, ptr, shape, dtype):
        return DevArray(shape, dtype, ptr=ptr, owner=False)

    def free(self):
        if self.owner and self.ptr:
            lib().lgn_device_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MappedHostArray:
    """pinned + mapped host memory (UVA zero-copy tier; host_alloc_space, Kernels.cu:57-64)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        h, d = C.c_void_p(), C.c_void_p()
        check(lib().lgn_host_alloc_mapped(C.byref(h), C.byref(d), C.c_int64(self.nbytes)), "lgn_host_alloc_mapped")
        self.host_ptr, self.ptr = h.value, d.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.host_ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    @classmethod
    def from_numpy(cls, a):
        m = cls(a.shape, a.dtype)
        m.array[...] = a
        return m

    def free(self):
        if self.host_ptr:
            self.array = None
            lib().lgn_host_free(C.c_void_p(self.host_ptr))
            self.host_ptr = None
