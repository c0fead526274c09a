"""One-process-per-GPU plumbing (torch.distributed): seed partitioning, the path's single collective
(sum of the presampling hotness histograms, replacing aggregate_access, GPUCache.cu:44-48,624-647) and
the one-off exchange of CUDA-IPC handles of the cache shards.  No compute happens here: device work
stays behind the C-ABI; these helpers only move small host objects and call all_reduce on a tensor
view of the library's histogram.  Works with backend "nccl" (GPU tensors) and "gloo" (CPU tests)."""
import numpy as np


def partition_seeds(ids, world, rank):
    """seed `tid` belongs to partition `tid % P`, file order kept (GPUGraphStore.cu:332-346)."""
    return ids[(ids % world) == rank]


def train_steps(n_train_mine, batch, dist=None):
    """train_step = (min_i n_train_i - 1) / B (CUDA_IPC_Service.cu:71-88): the minimum over ranks."""
    n = int(n_train_mine)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        import torch
        t = torch.tensor([n], dtype=torch.int64)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t.item())
    return (n - 1) // batch


class _DeviceView:
    """zero-copy int32 view of a device array for torch.as_tensor (counts stay < 2^31)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def allreduce_hotness(dist, hist, n=None, device=None):
    """in-place sum over ranks.  `hist` is a torch tensor (CPU for gloo, CUDA for nccl) or a device
    array object with .ptr owned by the library."""
    import torch
    if hasattr(hist, "ptr"):
        t = torch.as_tensor(_DeviceView(hist.ptr, n), device=device)
    else:
        t = hist
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def native_allreduce_u32(dist, ptr, n, stream=None):
    """in-place sum of a device u32[n] histogram over the ranks with the LIBRARY's own NCCL communicator
    (lgn_comm_*: ncclAllReduce inside liblegion_b200.so); torch.distributed only ships the 128-byte unique id."""
    import ctypes as C
    from ._lib import lib, check
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        check(lib().lgn_comm_unique_id(uid), "lgn_comm_unique_id")
    box = [bytes(uid) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
    comm = C.c_void_p()
    check(lib().lgn_comm_create(C.c_int32(rank), C.c_int32(world), uid, C.byref(comm)), "lgn_comm_create")
    try:
        check(lib().lgn_comm_allreduce_u32(comm, C.c_void_p(int(ptr)), C.c_int64(n), C.c_void_p(stream or 0)), "lgn_comm_allreduce_u32")
        check(lib().lgn_stream_synchronize(C.c_void_p(stream or 0)), "lgn_stream_synchronize")
    finally:
        lib().lgn_comm_destroy(comm)


def exchange_handles(dist, handle_bytes):
    """every rank contributes its 64-byte CUDA-IPC handle; returns the list indexed by rank."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    if world == 1:
        return [bytes(handle_bytes)]
    out = [None] * world
    dist.all_gather_object(out, bytes(handle_bytes))
    return out


def exchange_fds(dist, fd, tag="shard"):
    """every rank contributes one open file descriptor (the shareable handle of lgn_shared_alloc) and receives a
    duplicate of every other rank's: descriptors cannot travel through torch.distributed, so each rank serves its own
    over a Unix-domain socket (SCM_RIGHTS).  Returns the list indexed by rank (own entry = the descriptor passed in);
    the caller closes the received duplicates after importing them.  Single node only, like the NVLink clique."""
    import os
    import socket
    import tempfile
    import threading
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    if world == 1:
        return [fd]
    rank = dist.get_rank()
    token = [os.urandom(6).hex() if rank == 0 else None]
    dist.broadcast_object_list(token, src=0)
    path = lambda j: os.path.join(tempfile.gettempdir(), "lgn_%s_%s_%d.sock" % (tag, token[0], j))
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    if os.path.exists(path(rank)):
        os.unlink(path(rank))
    srv.bind(path(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _addr = srv.accept()
            with conn:
                conn.recv(4)                                     # the requester's rank (informational)
                socket.send_fds(conn, [b"f"], [fd])

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    dist.barrier()                                               # every listener is up
    out = [None] * world
    out[rank] = fd
    for j in range(world):
        if j == rank:
            continue
        with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
            c.connect(path(j))
            c.sendall(int(rank).to_bytes(4, "little"))
            _msg, fds, _flags, _addr = socket.recv_fds(c, 1, 1)
            if len(fds) != 1:
                raise RuntimeError("rank %d sent no descriptor" % j)
            out[j] = fds[0]
    th.join()
    dist.barrier()
    srv.close()
    os.unlink(path(rank))
    return out


def slot_of_rank(i, cap, kg):
    """global slot of hot rank i: GPU i % kg, row i / kg (InitPair, GPUCache.cu:103-108)."""
    return (i % kg) * cap + i // kg


def shard_ranks(cap, kg, j, n):
    """hot ranks stored on GPU j, in row order (FeatFillUp, GPUCache.cu:200-205)."""
    r = np.arange(cap, dtype=np.int64) * kg + j
    return r[r < n]


def capacity_for(n_cached, kg):
    return max(1, (int(n_cached) + kg - 1) // kg)


def compact_split(n_nodes, n_repl, cap, kg):
    """rows of the compact placement (lgn_place_compact) for a shard height `cap`: -> (n_repl, n_part) with the n_repl hottest
    ranks replicated on every GPU and the next n_part dealt over kg GPUs (rank q of the class -> GPU q % kg, row n_repl + q // kg)."""
    n_repl = min(int(n_repl), int(n_nodes))
    cached = min(int(n_nodes), n_repl + (int(cap) - n_repl) * int(kg))
    n_part = max(0, cached - n_repl)
    assert int(cap) >= n_repl + (n_part + kg - 1) // kg, "shard too low for the partitioned class"
    return n_repl, n_part
