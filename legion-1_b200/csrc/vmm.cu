// vmm.cu -- cache shards that other processes map through the CUDA virtual-memory-management API.
//
// Alternative to the legacy cudaIpc* handles of lgn_ipc_export / lgn_ipc_import for the cross-process peer
// shards of the one-process-per-GPU deployment: the owner creates the physical allocation with a POSIX file
// descriptor as its shareable handle, every peer process imports the descriptor, maps it at an address of
// its own and grants ITS device read/write access explicitly (cuMemSetAccess), which is how NCCL maps its
// NVLink buffers.  No reference counterpart: the reference is one process with all GPUs (GPUGraphStore.cu:145-168).
//
// The driver entry points are resolved at run time (cudaGetDriverEntryPoint) so that liblegion_b200.so keeps
// loading on machines without libcuda.so.1 (the CPU test suite loads it to check the exported symbols).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <mutex>
#include <unordered_map>

#include "context.h"

int lgn_cuda_fail(cudaError_t e, const char* what);
extern thread_local char g_lgn_cuda_err[256];

namespace {

struct Driver {
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    CUresult (*MemRelease)(CUmemGenericAllocationHandle);
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*MemAddressFree)(CUdeviceptr, size_t);
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*MemUnmap)(CUdeviceptr, size_t);
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    CUresult (*GetErrorString)(CUresult, const char**);
    bool ok;
};

Driver* driver()
{
    static Driver d;
    static std::once_flag once;
    std::call_once(once, []() {
        memset(&d, 0, sizeof(d));
        struct { const char* name; void** slot; } syms[] = {
            {"cuMemCreate", (void**)&d.MemCreate}, {"cuMemRelease", (void**)&d.MemRelease},
            {"cuMemAddressReserve", (void**)&d.MemAddressReserve}, {"cuMemAddressFree", (void**)&d.MemAddressFree},
            {"cuMemMap", (void**)&d.MemMap}, {"cuMemUnmap", (void**)&d.MemUnmap}, {"cuMemSetAccess", (void**)&d.MemSetAccess},
            {"cuMemExportToShareableHandle", (void**)&d.MemExportToShareableHandle},
            {"cuMemImportFromShareableHandle", (void**)&d.MemImportFromShareableHandle},
            {"cuMemGetAllocationGranularity", (void**)&d.MemGetAllocationGranularity}, {"cuGetErrorString", (void**)&d.GetErrorString}};
        d.ok = true;
        for (auto& s : syms) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.slot) {
                d.ok = false;
                cudaGetLastError();
            }
        }
    });
    return d.ok ? &d : nullptr;
}

int drv_fail(Driver* d, CUresult r, const char* what)
{
    const char* msg = nullptr;
    if (d && d->GetErrorString) d->GetErrorString(r, &msg);
    snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "%s: %s", what, msg ? msg : "driver error");
    return LGN_E_CUDA;
}
#define DRV(call, what)                                   \
    do {                                                  \
        CUresult r_ = (call);                             \
        if (r_ != CUDA_SUCCESS) { rc = drv_fail(d, r_, what); goto fail; } \
    } while (0)

struct Mapping { CUmemGenericAllocationHandle handle; size_t size; };
std::mutex g_mu;
std::unordered_map<void*, Mapping> g_maps;

CUmemAllocationProp device_prop(int device)
{
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

// Size and address alignment of shared shards.  Random row reads out of multi-GB peer shards are bound by the
// reach of the GPU's address translation, not by NVLink: physical size and virtual address
// are therefore rounded to LGN_VMM_ALIGN_MB (default 512 MB, the largest GPU page) so the driver can map the shard
// with its largest pages on the owner and on every importer.
size_t huge_granularity(size_t gran)
{
    const char* e = getenv("LGN_VMM_ALIGN_MB");
    size_t want = (size_t)(e ? atoll(e) : 512) << 20;
    if (want < gran) want = gran;
    return (want + gran - 1) / gran * gran;
}

// map `handle` (size bytes, already a multiple of the granularity) into this process and give `device` access
int map_for_device(Driver* d, CUmemGenericAllocationHandle handle, size_t size, size_t gran, int device, void** out)
{
    int rc = LGN_OK;
    CUdeviceptr va = 0;
    bool mapped = false;
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    DRV(d->MemAddressReserve(&va, size, gran, 0, 0), "cuMemAddressReserve");
    DRV(d->MemMap(va, size, 0, handle, 0), "cuMemMap");
    mapped = true;
    DRV(d->MemSetAccess(va, size, &acc, 1), "cuMemSetAccess");
    {
        std::lock_guard<std::mutex> g(g_mu);
        g_maps[(void*)va] = Mapping{handle, size};
    }
    *out = (void*)va;
    return LGN_OK;
fail:
    if (mapped) d->MemUnmap(va, size);
    if (va) d->MemAddressFree(va, size);
    return rc;
}

}  // namespace

extern "C" {

int lgn_shared_alloc(void** dev_ptr, int64_t bytes, int32_t* fd_out, int64_t* mapped_bytes)
{
    if (!dev_ptr || !fd_out || bytes <= 0) return LGN_E_ARG;
    int device = 0;
    cudaError_t ce = cudaGetDevice(&device);
    if (ce != cudaSuccess) return lgn_cuda_fail(ce, "cudaGetDevice");
    ce = cudaFree(0);                                   // make sure the primary context exists before driver calls
    if (ce != cudaSuccess) return lgn_cuda_fail(ce, "cudaFree(0)");
    Driver* d = driver();
    if (!d) { snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "CUDA driver VMM entry points unavailable"); return LGN_E_CUDA; }
    int rc = LGN_OK;
    CUmemAllocationProp prop = device_prop(device);
    CUmemGenericAllocationHandle handle = 0;
    bool created = false;
    size_t gran = 0, size = 0;
    int fd = -1;
    DRV(d->MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    gran = huge_granularity(gran);
    size = ((size_t)bytes + gran - 1) / gran * gran;
    DRV(d->MemCreate(&handle, size, &prop, 0), "cuMemCreate");
    created = true;
    DRV(d->MemExportToShareableHandle(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    rc = map_for_device(d, handle, size, gran, device, dev_ptr);
    if (rc) goto fail;
    *fd_out = fd;
    if (mapped_bytes) *mapped_bytes = (int64_t)size;
    return LGN_OK;
fail:
    if (fd >= 0) close(fd);
    if (created) d->MemRelease(handle);
    return rc;
}

int lgn_shared_import(int32_t fd, int64_t mapped_bytes, void** dev_ptr)
{
    if (!dev_ptr || fd < 0 || mapped_bytes <= 0) return LGN_E_ARG;
    int device = 0;
    cudaError_t ce = cudaGetDevice(&device);
    if (ce != cudaSuccess) return lgn_cuda_fail(ce, "cudaGetDevice");
    ce = cudaFree(0);
    if (ce != cudaSuccess) return lgn_cuda_fail(ce, "cudaFree(0)");
    Driver* d = driver();
    if (!d) { snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "CUDA driver VMM entry points unavailable"); return LGN_E_CUDA; }
    int rc = LGN_OK;
    CUmemAllocationProp prop = device_prop(device);
    CUmemGenericAllocationHandle handle = 0;
    bool imported = false;
    size_t gran = 0;
    DRV(d->MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    gran = huge_granularity(gran);
    DRV(d->MemImportFromShareableHandle(&handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
    imported = true;
    rc = map_for_device(d, handle, (size_t)mapped_bytes, gran, device, dev_ptr);   // access is granted to THIS process's device
    if (rc) goto fail;
    return LGN_OK;
fail:
    if (imported) d->MemRelease(handle);
    return rc;
}

int lgn_shared_free(void* dev_ptr)
{
    if (!dev_ptr) return LGN_E_ARG;
    Driver* d = driver();
    if (!d) return LGN_E_CUDA;
    Mapping m;
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_maps.find(dev_ptr);
        if (it == g_maps.end()) return LGN_E_ARG;
        m = it->second;
        g_maps.erase(it);
    }
    d->MemUnmap((CUdeviceptr)dev_ptr, m.size);
    d->MemAddressFree((CUdeviceptr)dev_ptr, m.size);
    d->MemRelease(m.handle);
    return LGN_OK;
}

}  // extern "C"
