"""ipc_service -- drop-in replacement of the reference's pybind11 extension of the same name
(pytorch_extension/ipc_service.cpp:86-93): initialize / finalize / get_steps / get_next /
get_block_size / synchronize, same argument meaning, same return layout, same wire format
(POSIX shm "simpleIPCshm", sem_r_/sem_w_<dev>_<pipe>, 7 CUDA-IPC buffers per slot).

Put this directory on PYTHONPATH and the unchanged trainers (legion_graphsage.py, legion_gcn.py,
lp_sage.py) `import ipc_service` as before.  The tensors returned by get_next alias server-owned
IPC memory (zero copy) and stay valid until synchronize(), like the reference's from_blob views
(ipc_cuda_kernel.cu:198-229).  Call torch.cuda.set_device(rank) first: everything is keyed on the
current CUDA device (ipc_cuda_kernel.cu:41,63).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_client = None
_h_nc = (C.c_int32 * 16)()
_h_ec = (C.c_int32 * 16)()
_TYPESTR = {torch.int32: "<i4", torch.float32: "<f4"}


def _load():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "_build", "liblegion_b200.so")
        if not os.path.exists(so):
            raise RuntimeError(f"{so} missing: build the library first (python -c 'import __graft_entry__ as g; g.build()')")
        _lib = C.CDLL(so)
        _lib.lgn_error_string.restype = C.c_char_p
        _lib.lgn_last_cuda_error.restype = C.c_char_p
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"ipc_service.{what}: {_load().lgn_error_string(rc).decode()} {_load().lgn_last_cuda_error().decode()}")


class _Blob:
    """zero-copy view of device memory for torch.as_tensor (the reference uses torch::from_blob)."""

    def __init__(self, ptr, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, dtype, device):
    n = 1
    for s in shape:
        n *= s
    if n == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    return torch.as_tensor(_Blob(ptr, shape, dtype), device=device)


def initialize():
    """InitializeIPC (ipc_service.cpp:14-17): open shm, the 2x7 IPC handles of this device, the semaphores."""
    global _client
    dev = torch.cuda.current_device()
    h = C.c_void_p()
    _check(_load().lgn_ipc_client_open(C.c_int32(dev), C.byref(h)), "initialize")
    _client = h


def finalize():
    global _client
    if _client is not None:
        _load().lgn_ipc_client_close(_client)
        _client = None


def get_steps():
    s = (C.c_int32 * 3)()
    _check(_load().lgn_ipc_client_steps(_client, s), "get_steps")
    return [int(s[0]), int(s[1]), int(s[2])]


def _next():
    ptrs = (C.c_void_p * 7)()
    _check(_load().lgn_ipc_client_next(_client, ptrs, _h_nc, _h_ec), "get_next")
    return ptrs, torch.device("cuda", torch.cuda.current_device())


def get_next(feature_dim):
    """[ids, features, labels, block1_src, block1_dst, block2_src, block2_dst] (ipc_service.cpp:43-58):
    block1 = COO prefix ec[4] (hop-1 and hop-2 edges), block2 = prefix ec[3] (hop-1 edges)."""
    ptrs, dev = _next()
    nc, ec = _h_nc, _h_ec
    ids = _view(ptrs[0], (nc[9],), torch.int32, dev)
    feats = _view(ptrs[1], (nc[9], feature_dim), torch.float32, dev)
    labels = _view(ptrs[2], (nc[5],), torch.int32, dev)
    return [ids, feats, labels,
            _view(ptrs[3], (ec[4],), torch.int32, dev), _view(ptrs[4], (ec[4],), torch.int32, dev),
            _view(ptrs[3], (ec[3],), torch.int32, dev), _view(ptrs[4], (ec[3],), torch.int32, dev)]


def get_block_size():
    """[block1 #src, block1 #dst, block2 #src, block2 #dst] = [nc9, nc7, nc7, nc5] (ipc_service.cpp:60-72)."""
    nc = _h_nc
    return [int(nc[9]), int(nc[7]), int(nc[7]), int(nc[5])]


def synchronize():
    """release the slot to the server and move to the other one (Post, ipc_cuda_kernel.cu:102-106)."""
    _check(_load().lgn_ipc_client_release(_client), "synchronize")


# ---- additive k-hop API (the reference's consumer is hard-wired to two hops) ----
def get_next_k(feature_dim, n_hops):
    """[ids, features, labels, [(src, dst) of GNN layer 1 .. layer n_hops]]: layer l consumes the COO prefix
    through hop n_hops-l+1, so layer 1 sees every sampled edge and the last layer only hop 1's."""
    ptrs, dev = _next()
    nc, ec = _h_nc, _h_ec
    total = nc[7 + 2 * (n_hops - 1)] if n_hops > 0 else nc[4]
    ids = _view(ptrs[0], (total,), torch.int32, dev)
    feats = _view(ptrs[1], (total, feature_dim), torch.float32, dev)
    labels = _view(ptrs[2], (nc[4],), torch.int32, dev)
    blocks = []
    for layer in range(n_hops):
        e = ec[3 + (n_hops - 1 - layer)]
        blocks.append((_view(ptrs[3], (e,), torch.int32, dev), _view(ptrs[4], (e,), torch.int32, dev)))
    return [ids, feats, labels, blocks]


def get_block_sizes_k(n_hops):
    """[(#src, #dst)] per GNN layer, outermost first: dst nodes are always a prefix of src nodes."""
    nc = _h_nc
    out = []
    for layer in range(n_hops):
        h = n_hops - 1 - layer            # deepest hop this layer consumes (0-based)
        out.append((int(nc[7 + 2 * h]), int(nc[5 + 2 * h])))
    return out
