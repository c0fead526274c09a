// ipc.cu -- server<->trainer wire format, both ends.
//
// Byte-compatible with the reference so its trainers (and its own ipc_service extension)
// keep working: POSIX shm "simpleIPCshm" holding {int32 steps[3]; cudaIpcMemHandle_t
// memHandle[8][2][7]} (CUDA_IPC_Service.cu:34-37, ipc_cuda_kernel.cu:30-33), named
// semaphores sem_r_<dev>_<pipe> / sem_w_<dev>_<pipe> created with value 0
// (CUDA_IPC_Service.cu:189-199); handle index 0 ids, 1 features, 2 labels, 3 agg_src,
// 4 agg_dst, 5 node_counter, 6 edge_counter (CUDA_IPC_Service.cu:169-175, 209).
#include <errno.h>
#include <fcntl.h>
#include <semaphore.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <new>

#include "context.h"

int lgn_cuda_fail(cudaError_t e, const char* what);
#define CK(x)                                                 \
    do {                                                      \
        cudaError_t e_ = (x);                                 \
        if (e_ != cudaSuccess) return lgn_cuda_fail(e_, #x);  \
    } while (0)

namespace {

constexpr int MEMORY_USAGE = 7;
struct ShmLayout {
    int32_t steps[3];
    cudaIpcMemHandle_t mem[LGN_MAX_PARTS][LGN_PIPELINE_DEPTH][MEMORY_USAGE];
};
static_assert(sizeof(ShmLayout) == 3 * 4 + 8 * 2 * 7 * 64, "wire format: 7180 bytes, no padding");
const char* kShmName = "simpleIPCshm";

int map_shm(bool create, int* fd_out, ShmLayout** out)
{
    int fd = shm_open(kShmName, create ? (O_RDWR | O_CREAT) : O_RDWR, 0777);   // helper_multiprocess.cpp:30
    if (fd < 0) return LGN_E_SYS;
    if (create && ftruncate(fd, sizeof(ShmLayout)) != 0) { close(fd); return LGN_E_SYS; }
    void* p = mmap(0, sizeof(ShmLayout), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) { close(fd); return LGN_E_SYS; }
    *fd_out = fd;
    *out = (ShmLayout*)p;
    return LGN_OK;
}

void sem_name(char* buf, size_t n, const char* pre, int dev, int pipe) { snprintf(buf, n, "%s%d_%d", pre, dev, pipe); }

}  // namespace

struct lgn_ipc_server {
    int n_dev, fd;
    ShmLayout* shm;
    sem_t* semr[LGN_MAX_PARTS][LGN_PIPELINE_DEPTH];
    sem_t* semw[LGN_MAX_PARTS][LGN_PIPELINE_DEPTH];
};

struct lgn_ipc_client {
    int dev, fd, pipe;
    ShmLayout* shm;
    void* ptr[LGN_PIPELINE_DEPTH][MEMORY_USAGE];
    sem_t* semr[LGN_PIPELINE_DEPTH];
    sem_t* semw[LGN_PIPELINE_DEPTH];
};

extern "C" {

int lgn_ipc_server_create(int32_t n_devices, const int32_t steps[3], lgn_ipc_server** out)
{
    if (n_devices <= 0 || n_devices > LGN_MAX_PARTS || !steps || !out) return LGN_E_ARG;
    lgn_ipc_server* s = new (std::nothrow) lgn_ipc_server();
    if (!s) return LGN_E_SYS;
    memset(s, 0, sizeof(*s));
    s->n_dev = n_devices;
    int rc = map_shm(true, &s->fd, &s->shm);
    if (rc) { delete s; return rc; }
    memset(s->shm, 0, sizeof(ShmLayout));                                        // CUDA_IPC_Service.cu:50
    for (int i = 0; i < 3; i++) s->shm->steps[i] = steps[i];                     // :129-131
    char name[64];
    for (int d = 0; d < n_devices; d++)
        for (int p = 0; p < LGN_PIPELINE_DEPTH; p++) {
            // stale semaphores of a crashed run would carry old counts: start from a clean slate
            sem_name(name, sizeof(name), "sem_r_", d, p); sem_unlink(name);
            s->semr[d][p] = sem_open(name, O_CREAT | O_RDWR, 0666, 0);
            sem_name(name, sizeof(name), "sem_w_", d, p); sem_unlink(name);
            s->semw[d][p] = sem_open(name, O_CREAT | O_RDWR, 0666, 0);
            if (s->semr[d][p] == SEM_FAILED || s->semw[d][p] == SEM_FAILED) { lgn_ipc_server_destroy(s); return LGN_E_SYS; }
        }
    *out = s;
    return LGN_OK;
}

int lgn_ipc_server_publish(lgn_ipc_server* s, int32_t device, lgn_ctx* c, int32_t with_features)
{
    if (!s || !c || device < 0 || device >= s->n_dev) return LGN_E_ARG;
    CK(cudaSetDevice(c->cfg.device));
    for (int p = 0; p < LGN_PIPELINE_DEPTH; p++) {
        const lgn::Pipe& pp = c->pipe[p];
        void* bufs[MEMORY_USAGE] = {pp.ids, with_features ? (void*)pp.features : nullptr, pp.labels, pp.agg_src_off, pp.agg_dst_off, pp.nc, pp.ec};
        for (int k = 0; k < MEMORY_USAGE; k++) {
            if (!bufs[k]) continue;   // features are published after presampling (Server.cu:273-282)
            cudaIpcMemHandle_t h;
            CK(cudaIpcGetMemHandle(&h, bufs[k]));
            memcpy((void*)&s->shm->mem[device][p][k], &h, sizeof(h));
        }
    }
    return LGN_OK;
}

int lgn_ipc_server_wait(lgn_ipc_server* s, int32_t device, int32_t pipe)
{
    if (!s || device < 0 || device >= s->n_dev || pipe < 0 || pipe >= LGN_PIPELINE_DEPTH) return LGN_E_ARG;
    while (sem_wait(s->semr[device][pipe]) != 0) if (errno != EINTR) return LGN_E_SYS;
    return LGN_OK;
}

int lgn_ipc_server_post(lgn_ipc_server* s, int32_t device, int32_t pipe)
{
    if (!s || device < 0 || device >= s->n_dev || pipe < 0 || pipe >= LGN_PIPELINE_DEPTH) return LGN_E_ARG;
    return sem_post(s->semw[device][pipe]) == 0 ? LGN_OK : LGN_E_SYS;
}

int lgn_ipc_server_destroy(lgn_ipc_server* s)
{
    if (!s) return LGN_E_ARG;
    char name[64];
    for (int d = 0; d < s->n_dev; d++)
        for (int p = 0; p < LGN_PIPELINE_DEPTH; p++) {                           // CUDA_IPC_Service.cu:300-312
            if (s->semr[d][p] && s->semr[d][p] != SEM_FAILED) sem_close(s->semr[d][p]);
            if (s->semw[d][p] && s->semw[d][p] != SEM_FAILED) sem_close(s->semw[d][p]);
            sem_name(name, sizeof(name), "sem_r_", d, p); sem_unlink(name);
            sem_name(name, sizeof(name), "sem_w_", d, p); sem_unlink(name);
        }
    if (s->shm) { munmap(s->shm, sizeof(ShmLayout)); close(s->fd); }
    shm_unlink(kShmName);
    delete s;
    return LGN_OK;
}

int lgn_ipc_client_open(int32_t device, lgn_ipc_client** out)
{
    if (device < 0 || device >= LGN_MAX_PARTS || !out) return LGN_E_ARG;
    lgn_ipc_client* c = new (std::nothrow) lgn_ipc_client();
    if (!c) return LGN_E_SYS;
    memset(c, 0, sizeof(*c));
    c->dev = device;
    int rc = map_shm(false, &c->fd, &c->shm);
    if (rc) { delete c; return rc; }
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) { lgn_ipc_client_close(c); return lgn_cuda_fail(ce, "cudaSetDevice"); }
    char name[64];
    for (int p = 0; p < LGN_PIPELINE_DEPTH; p++) {                                // ipc_cuda_kernel.cu:62-92
        for (int k = 0; k < MEMORY_USAGE; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (void*)&c->shm->mem[device][p][k], sizeof(h));
            static const cudaIpcMemHandle_t zero = {};
            if (memcmp(&h, &zero, sizeof(h)) == 0) continue;   // not published (yet)
            ce = cudaIpcOpenMemHandle(&c->ptr[p][k], h, cudaIpcMemLazyEnablePeerAccess);
            if (ce != cudaSuccess) { c->ptr[p][k] = nullptr; lgn_ipc_client_close(c); return lgn_cuda_fail(ce, "cudaIpcOpenMemHandle"); }
        }
        sem_name(name, sizeof(name), "sem_r_", device, p);
        c->semr[p] = sem_open(name, O_CREAT | O_RDWR, 0666, 0);
        sem_name(name, sizeof(name), "sem_w_", device, p);
        c->semw[p] = sem_open(name, O_CREAT | O_RDWR, 0666, 0);
        if (c->semr[p] == SEM_FAILED || c->semw[p] == SEM_FAILED) { lgn_ipc_client_close(c); return LGN_E_SYS; }
        sem_post(c->semr[p]);                                                      // both slots start free (:91)
    }
    *out = c;
    return LGN_OK;
}

int lgn_ipc_client_steps(lgn_ipc_client* c, int32_t steps[3])
{
    if (!c || !steps) return LGN_E_ARG;
    for (int i = 0; i < 3; i++) steps[i] = c->shm->steps[i];
    return LGN_OK;
}

int lgn_ipc_client_next(lgn_ipc_client* c, void* ptrs[7], int32_t nc[16], int32_t ec[16])
{
    if (!c || !ptrs || !nc || !ec) return LGN_E_ARG;
    while (sem_wait(c->semw[c->pipe]) != 0) if (errno != EINTR) return LGN_E_SYS;   // Wait(), :97-100
    for (int k = 0; k < MEMORY_USAGE; k++) ptrs[k] = c->ptr[c->pipe][k];
    CK(cudaMemcpy(nc, c->ptr[c->pipe][5], 64, cudaMemcpyDeviceToHost));             // :192-193
    CK(cudaMemcpy(ec, c->ptr[c->pipe][6], 64, cudaMemcpyDeviceToHost));
    return LGN_OK;
}

int lgn_ipc_client_release(lgn_ipc_client* c)
{
    if (!c) return LGN_E_ARG;
    if (sem_post(c->semr[c->pipe]) != 0) return LGN_E_SYS;                           // Post(), :102-106
    c->pipe = (c->pipe + 1) % LGN_PIPELINE_DEPTH;
    return LGN_OK;
}

int lgn_ipc_client_close(lgn_ipc_client* c)
{
    if (!c) return LGN_E_ARG;
    for (int p = 0; p < LGN_PIPELINE_DEPTH; p++) {
        for (int k = 0; k < MEMORY_USAGE; k++) if (c->ptr[p][k]) cudaIpcCloseMemHandle(c->ptr[p][k]);
        if (c->semr[p] && c->semr[p] != SEM_FAILED) sem_close(c->semr[p]);
        if (c->semw[p] && c->semw[p] != SEM_FAILED) sem_close(c->semw[p]);
    }
    cudaGetLastError();
    if (c->shm) { munmap(c->shm, sizeof(ShmLayout)); close(c->fd); }
    delete c;
    return LGN_OK;
}

}  // extern "C"
