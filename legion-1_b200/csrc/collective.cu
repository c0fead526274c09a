// collective.cu -- the path's ONE collective: the sum of the per-GPU presampling hotness histograms before cache
// planning.  The reference has the clique leader read its peers' arrays over P2P
// (aggregate_access, GPUCache.cu:44-48, 624-647); here it is an NCCL all-reduce of u32[N] over NVLink / NVSwitch, so
// every GPU ends up with the clique's histogram and can order and fill its own shard.
//
// Two shapes: one process driving all GPUs of the clique (the `legion` server: lgn_allreduce_u32_devices) and one
// process per GPU (bench.py / a torch.distributed job: lgn_comm_* with the 128-byte NCCL unique id exchanged by the
// caller).  NCCL is resolved with dlopen at first use: liblegion_b200.so keeps loading on machines without it, and a
// process that already loaded a libnccl.so.2 (PyTorch bundles one) shares that copy instead of pulling in a second.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "context.h"

int lgn_cuda_fail(cudaError_t e, const char* what);
extern thread_local char g_lgn_cuda_err[256];

namespace {

struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

Nccl* nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, []() {
        memset(&n, 0, sizeof(n));
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        struct { const char* name; void** slot; } syms[] = {
            {"ncclGetUniqueId", (void**)&n.GetUniqueId}, {"ncclCommInitRank", (void**)&n.CommInitRank},
            {"ncclCommInitAll", (void**)&n.CommInitAll}, {"ncclCommDestroy", (void**)&n.CommDestroy},
            {"ncclAllReduce", (void**)&n.AllReduce}, {"ncclGroupStart", (void**)&n.GroupStart},
            {"ncclGroupEnd", (void**)&n.GroupEnd}, {"ncclGetErrorString", (void**)&n.GetErrorString}};
        n.ok = true;
        for (auto& s : syms) { *s.slot = dlsym(h, s.name); if (!*s.slot) n.ok = false; }
    });
    return n.ok ? &n : nullptr;
}

int nccl_fail(Nccl* n, ncclResult_t r, const char* what)
{
    snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "%s: %s", what, n && n->GetErrorString ? n->GetErrorString(r) : "NCCL error");
    return LGN_E_CUDA;
}
#define NC(call, what)                                          \
    do {                                                        \
        ncclResult_t r_ = (call);                               \
        if (r_ != ncclSuccess) return nccl_fail(n, r_, what);   \
    } while (0)

}  // namespace

struct lgn_comm {
    ncclComm_t comm;
    int rank, world;
};

extern "C" {

int lgn_comm_available(void) { return nccl() ? 1 : 0; }

int lgn_comm_unique_id(uint8_t id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "the C-ABI carries NCCL's unique id as 128 opaque bytes");
    Nccl* n = nccl();
    if (!n) { snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "libnccl.so.2 not found"); return LGN_E_SYS; }
    if (!id) return LGN_E_ARG;
    ncclUniqueId u;
    NC(n->GetUniqueId(&u), "ncclGetUniqueId");
    memcpy(id, &u, 128);
    return LGN_OK;
}

int lgn_comm_create(int32_t rank, int32_t world, const uint8_t id[128], lgn_comm** out)
{
    Nccl* n = nccl();
    if (!n) { snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "libnccl.so.2 not found"); return LGN_E_SYS; }
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return LGN_E_ARG;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    lgn_comm* c = new lgn_comm();
    c->rank = rank; c->world = world; c->comm = nullptr;
    ncclResult_t r = n->CommInitRank(&c->comm, world, u, rank);      // the current device of the calling thread joins
    if (r != ncclSuccess) { delete c; return nccl_fail(n, r, "ncclCommInitRank"); }
    *out = c;
    return LGN_OK;
}

// in-place sum of u32[n] over the communicator's ranks, on `stream`
int lgn_comm_allreduce_u32(lgn_comm* c, uint32_t* buf, int64_t count, void* stream)
{
    Nccl* n = nccl();
    if (!n || !c || !buf || count < 0) return LGN_E_ARG;
    NC(n->AllReduce(buf, buf, (size_t)count, ncclUint32, ncclSum, c->comm, (cudaStream_t)stream), "ncclAllReduce");
    return LGN_OK;
}

int lgn_comm_destroy(lgn_comm* c)
{
    Nccl* n = nccl();
    if (!c) return LGN_E_ARG;
    if (n && c->comm) n->CommDestroy(c->comm);
    delete c;
    return LGN_OK;
}

// one process, several GPUs: bufs[i] lives on devices[i]; every buffer ends up holding the element-wise sum.
// Blocks until the result is complete on every device.
int lgn_allreduce_u32_devices(int32_t n_dev, const int32_t* devices, uint32_t* const* bufs, int64_t count)
{
    Nccl* n = nccl();
    if (!n) { snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "libnccl.so.2 not found"); return LGN_E_SYS; }
    if (n_dev < 1 || n_dev > LGN_MAX_PARTS || !devices || !bufs || count < 0) return LGN_E_ARG;
    if (n_dev == 1) return LGN_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    std::vector<ncclComm_t> comms(n_dev);
    std::vector<int> devs(devices, devices + n_dev);
    NC(n->CommInitAll(comms.data(), n_dev, devs.data()), "ncclCommInitAll");
    std::vector<cudaStream_t> streams(n_dev);
    int rc = LGN_OK;
    for (int i = 0; i < n_dev; i++) { cudaSetDevice(devs[i]); cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking); }
    ncclResult_t r = n->GroupStart();
    for (int i = 0; i < n_dev && r == ncclSuccess; i++)
        r = n->AllReduce(bufs[i], bufs[i], (size_t)count, ncclUint32, ncclSum, comms[i], streams[i]);
    if (r == ncclSuccess) r = n->GroupEnd();
    if (r != ncclSuccess) rc = nccl_fail(n, r, "ncclAllReduce (group)");
    for (int i = 0; i < n_dev; i++) {
        cudaSetDevice(devs[i]);
        cudaError_t e = cudaStreamSynchronize(streams[i]);
        if (e != cudaSuccess && rc == LGN_OK) rc = lgn_cuda_fail(e, "cudaStreamSynchronize");
        cudaStreamDestroy(streams[i]);
        n->CommDestroy(comms[i]);
    }
    cudaSetDevice(cur);
    return rc;
}

}  // extern "C"
