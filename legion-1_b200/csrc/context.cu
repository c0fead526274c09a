// context.cu -- C-ABI entry points: memory helpers, per-GPU context, storage binding and
// the operator calls (include/legion_b200.h).  Host logic mirrors GPURunner
// (Server.cu:167-364) and the extern "C" wrappers of Kernels.cu without their blocking
// cudaMemcpy / malloc on the hot loop (Kernels.cu:605-625, GPUCache.cu:394-395).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "context.h"

thread_local char g_lgn_cuda_err[256] = "";

int lgn_cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_lgn_cuda_err, sizeof(g_lgn_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return LGN_E_CUDA;
}
#define CK(x)                                                 \
    do {                                                      \
        cudaError_t e_ = (x);                                 \
        if (e_ != cudaSuccess) return lgn_cuda_fail(e_, #x);  \
    } while (0)

using namespace lgn;

template <typename T>
static cudaError_t malloc_pages(T** p, size_t n)
{
    const size_t page = (size_t)2 << 20;
    if (n == 0) n = 1;
    if (n >= page / 2) n = (n + page - 1) / page * page;
    return cudaMalloc((void**)p, n);
}

extern "C" {

const char* lgn_error_string(int code)
{
    switch (code) {
        case LGN_OK: return "ok";
        case LGN_E_ARG: return "invalid argument";
        case LGN_E_CUDA: return "CUDA error";
        case LGN_E_STATE: return "invalid call order / missing binding";
        case LGN_E_CAPACITY: return "batch buffer capacity exceeded";
        case LGN_E_SYS: return "system (shm/semaphore) error";
    }
    return "unknown";
}
const char* lgn_last_cuda_error(void) { return g_lgn_cuda_err; }
int lgn_version(void) { return 100; }

// ---------------------------------------------------------------- memory
int lgn_set_device(int32_t d) { CK(cudaSetDevice(d)); return LGN_OK; }
int lgn_get_device(int32_t* d) { int x = 0; CK(cudaGetDevice(&x)); *d = x; return LGN_OK; }
int lgn_device_count(int32_t* n) { int x = 0; CK(cudaGetDeviceCount(&x)); *n = x; return LGN_OK; }
// Any allocation of this library may end up mapped by a peer (cache shards over cudaIpc* handles or direct P2P).  A
// peer mapping of an allocation whose SIZE is not a multiple of the 2 MiB GPU page is built from small pages: random
// 512-byte row reads over 7 x 2.56 GB of such mappings ran at 26 GB/s per GPU instead of 585 GB/s on 8xB200, the
// translation misses stalling every other kernel on the GPU too (profiles/r02a_n8_collapse_diag_page_size.txt).
// cudaMalloc hands out whole 2 MiB pages for large requests anyway, so rounding the request costs no memory.
int lgn_device_alloc(void** p, int64_t bytes)
{
    if (!p || bytes < 0) return LGN_E_ARG;
    CK(malloc_pages(p, (size_t)bytes));
    return LGN_OK;
}
int lgn_device_free(void* p) { CK(cudaFree(p)); return LGN_OK; }
int lgn_host_alloc_mapped(void** host, void** dev, int64_t bytes)
{
    if (!host || !dev || bytes < 0) return LGN_E_ARG;
    CK(cudaHostAlloc(host, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocMapped | cudaHostAllocPortable));   // Kernels.cu:57-64
    CK(cudaHostGetDevicePointer(dev, *host, 0));
    return LGN_OK;
}
int lgn_host_free(void* host) { CK(cudaFreeHost(host)); return LGN_OK; }
int lgn_copy_h2d(void* d, const void* h, int64_t bytes) { CK(cudaMemcpy(d, h, (size_t)bytes, cudaMemcpyHostToDevice)); return LGN_OK; }
int lgn_copy_d2h(void* h, const void* d, int64_t bytes) { CK(cudaMemcpy(h, d, (size_t)bytes, cudaMemcpyDeviceToHost)); return LGN_OK; }
int lgn_memset_d(void* d, int v, int64_t bytes) { CK(cudaMemset(d, v, (size_t)bytes)); return LGN_OK; }
int lgn_copy_d2d(void* d, const void* s, int64_t bytes) { CK(cudaMemcpy(d, s, (size_t)bytes, cudaMemcpyDefault)); return LGN_OK; }
int lgn_device_synchronize(void) { CK(cudaDeviceSynchronize()); return LGN_OK; }
int lgn_copy_async(void* d, const void* s, int64_t bytes, void* stream)
{
    if (!d || !s || bytes < 0) return LGN_E_ARG;
    CK(cudaMemcpyAsync(d, s, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return LGN_OK;
}

int lgn_stream_create(void** stream, int32_t high_priority)
{
    if (!stream) return LGN_E_ARG;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    if (const char* lp = getenv("LGN_LANE_PRIO")) high_priority = lp[0] == 'h';     // experiment knob
    CK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
    *stream = (void*)s;
    return LGN_OK;
}
int lgn_stream_destroy(void* stream) { CK(cudaStreamDestroy((cudaStream_t)stream)); return LGN_OK; }
int lgn_stream_synchronize(void* stream) { CK(cudaStreamSynchronize((cudaStream_t)stream)); return LGN_OK; }

int lgn_enable_peer_access(int32_t n)
{
    int cur = 0;
    CK(cudaGetDevice(&cur));
    for (int i = 0; i < n; i++) {
        CK(cudaSetDevice(i));
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int ok = 0;
            CK(cudaDeviceCanAccessPeer(&ok, i, j));
            if (ok) {
                cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return lgn_cuda_fail(e, "cudaDeviceEnablePeerAccess");
            }
        }
    }
    CK(cudaSetDevice(cur));
    return LGN_OK;
}

int lgn_ipc_export(void* p, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "wire format assumes 64-byte handles");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle, &h, 64);
    return LGN_OK;
}
int lgn_ipc_import(const uint8_t handle[64], void** p)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    return LGN_OK;
}
int lgn_ipc_close(void* p) { CK(cudaIpcCloseMemHandle(p)); return LGN_OK; }

// ---------------------------------------------------------------- context
static void invalidate_graphs_impl(lgn_ctx* c);
static inline void invalidate_graphs(lgn_ctx* c) { invalidate_graphs_impl(c); }
static int fill_i32(int32_t* p, int32_t v, long long n, cudaStream_t s);

__global__ void k_fill_i32(int32_t* p, int32_t v, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
static int fill_i32(int32_t* p, int32_t v, long long n, cudaStream_t s)
{
    k_fill_i32<<<1024, 256, 0, s>>>(p, v, n);
    CK(cudaGetLastError());
    return LGN_OK;
}

static int create_impl(lgn_ctx* c, const lgn_config* cfg, long long cap, long long max_slots);

int lgn_create(const lgn_config* cfg, lgn_ctx** out)
{
    if (!cfg || !out) return LGN_E_ARG;
    *out = nullptr;
    if (cfg->n_nodes <= 0 || cfg->n_nodes > 0x7fffffffLL || cfg->feat_dim < 0 || cfg->batch_size <= 0) return LGN_E_ARG;
    if (cfg->n_hops < 0 || cfg->n_hops > LGN_MAX_HOPS || cfg->part < 0 || cfg->part >= LGN_MAX_PARTS) return LGN_E_ARG;
    if (cfg->rng_mode != LGN_RNG_MINSTD && cfg->rng_mode != LGN_RNG_PHILOX) return LGN_E_ARG;
    if (cfg->n_lanes < 0 || cfg->n_lanes > LGN_MAX_LANES) return LGN_E_ARG;
    long long cap = cfg->batch_size, cur = cfg->batch_size, max_slots = 1;
    for (int h = 0; h < cfg->n_hops; h++) {             // Server.cu:184-196
        if (cfg->fanout[h] <= 0 || cfg->fanout[h] > 256) return LGN_E_ARG;
        cur *= cfg->fanout[h];
        cap += cur;
        if (cur > max_slots) max_slots = cur;
        if (cap >= lgn::CAND || cur >= lgn::CAND) return LGN_E_ARG;   // local indices and slot indices must fit the 24-bit payload of a dedup value
    }
    CK(cudaSetDevice(cfg->device));
    lgn_ctx* c = new (std::nothrow) lgn_ctx();
    if (!c) return LGN_E_SYS;
    memset(c, 0, sizeof(*c));
    const int rc = create_impl(c, cfg, cap, max_slots);
    if (rc != LGN_OK) {          // release whatever was allocated before the failure; keep the CUDA error text
        char keep[sizeof(g_lgn_cuda_err)];
        memcpy(keep, g_lgn_cuda_err, sizeof(keep));
        lgn_destroy(c);
        memcpy(g_lgn_cuda_err, keep, sizeof(keep));
        return rc;
    }
    *out = c;
    return LGN_OK;
}

static int create_impl(lgn_ctx* c, const lgn_config* cfg, long long cap, long long max_slots)
{
    c->cfg = *cfg;
    c->capacity = cap;
    c->max_slots = max_slots;
    c->max_rows = cfg->max_feature_rows > 0 ? cfg->max_feature_rows : cap;
    c->n_lanes = cfg->n_lanes > 0 ? cfg->n_lanes : LGN_PIPELINE_DEPTH;
    CK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    {   // dedup layout: a batch-sized hash table (L2-resident for any N) unless the direct map itself is small
        const char* dm = getenv("LGN_DEDUP");
        const bool small_map = (size_t)cfg->n_nodes * 4 <= ((size_t)32 << 20);   // measured: direct wins at 9.8 MB (C2), hash at 444 MB (C3)
        c->dedup_hash = dm ? (dm[0] == 'h') : !small_map;
        uint32_t bits = 10;
        while (((long long)1 << bits) < 2 * cap && bits < 30) bits++;
        c->dedup_bits_max = bits;
    }
    {   // tiles of the widest hop (k_sample: sample_items_per_tile() frontier items each)
        long long items = cfg->batch_size, widest = cfg->batch_size;
        for (int h = 1; h < cfg->n_hops; h++) { items *= cfg->fanout[h - 1]; if (items > widest) widest = items; }
        c->max_tiles = (widest + lgn::sample_items_per_tile() - 1) / lgn::sample_items_per_tile() + 1;
    }
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int p = 0; p < c->n_lanes; p++) {      // CUDA_IPC_Service.cu:140-215, Server.cu:217-231
        lgn::Pipe& pp = c->pipe[p];
        CK(malloc_pages(&pp.ids, cap * sizeof(int32_t)));
        CK(malloc_pages(&pp.labels, (size_t)cfg->batch_size * sizeof(int32_t)));
        CK(malloc_pages(&pp.agg_src_off, cap * sizeof(int32_t)));
        CK(malloc_pages(&pp.agg_dst_off, cap * sizeof(int32_t)));
        CK(malloc_pages(&pp.nc, 16 * sizeof(int32_t)));
        CK(malloc_pages(&pp.ec, 16 * sizeof(int32_t)));
        CK(cudaMemset(pp.nc, 0, 64));
        CK(cudaMemset(pp.ec, 0, 64));
        if (cfg->feat_dim > 0) CK(malloc_pages(&pp.features, (size_t)c->max_rows * cfg->feat_dim * sizeof(float)));
        if (c->dedup_hash) {
            const size_t entries = (size_t)1 << c->dedup_bits_max;
            CK(malloc_pages(&pp.dedup_tab, entries * sizeof(unsigned long long)));
            CK(cudaMemset(pp.dedup_tab, 0xFF, entries * sizeof(unsigned long long)));
            CK(malloc_pages(&pp.seed_h, (size_t)cfg->batch_size * sizeof(int32_t)));
            CK(malloc_pages(&pp.draw_key, (max_slots + 16) * sizeof(int32_t)));
            pp.dedup.map = nullptr; pp.dedup.tab = pp.dedup_tab; pp.dedup.bits = c->dedup_bits_max;
        } else {
            CK(malloc_pages(&pp.slot_map, (size_t)cfg->n_nodes * sizeof(int32_t)));
            pp.dedup.map = pp.slot_map; pp.dedup.tab = nullptr; pp.dedup.bits = 0;
        }
        CK(malloc_pages(&pp.agg_src_ids, cap * sizeof(int32_t)));
        CK(malloc_pages(&pp.agg_dst_ids, cap * sizeof(int32_t)));
        CK(malloc_pages(&pp.draw_h, (max_slots + 16) * sizeof(int32_t)));
        CK(malloc_pages(&pp.draw_s, (max_slots + 16) * sizeof(uint16_t)));
        CK(malloc_pages(&pp.draw_v, (max_slots + 16) * sizeof(int32_t)));
        CK(malloc_pages(&pp.tile_n, c->max_tiles * sizeof(int32_t)));
        CK(malloc_pages(&pp.tile_new, c->max_tiles * sizeof(int32_t)));
        CK(malloc_pages(&pp.super_e, (c->max_tiles / 64 + 2) * sizeof(int32_t)));
        CK(malloc_pages(&pp.super_n, (c->max_tiles / 64 + 2) * sizeof(int32_t)));
        CK(malloc_pages(&pp.state, sizeof(lgn::BatchState)));
        CK(cudaMemset(pp.state, 0, sizeof(lgn::BatchState)));
        {
            const int32_t dbg[2] = {(int32_t)(max_slots + 16), (int32_t)cap};
            CK(cudaMemcpy(&pp.state->dbg_max_slots, dbg, sizeof(dbg), cudaMemcpyHostToDevice));
        }
        CK(malloc_pages(&pp.seed_stage, (size_t)cfg->batch_size * 2 * sizeof(int32_t)));
        if (!c->dedup_hash) {
            int rc = fill_i32(pp.slot_map, lgn::EMPTY, cfg->n_nodes, 0);
            if (rc) return rc;
        }
        // gathers run at the lowest priority so the latency-bound sampling kernels get SM slots first
        // (LGN_GATHER_PRIO=hi: experiment knob)
        const char* gp = getenv("LGN_GATHER_PRIO");
        CK(cudaStreamCreateWithPriority(&pp.gather_stream, cudaStreamNonBlocking, gp && gp[0] == 'h' ? prio_hi : prio_lo));
        for (int i = 0; i < LGN_MAX_HOPS + 2; i++) CK(cudaEventCreateWithFlags(&pp.ev_hop[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.ev_end, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.ev_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&pp.ev_join, cudaEventDisableTiming));
    }
    if (cfg->enable_hotness) {                                                   // GPUCache.cu:256-261 (u64 there)
        CK(malloc_pages(&c->node_hotness, (size_t)cfg->n_nodes * sizeof(uint32_t)));
        CK(malloc_pages(&c->topo_hotness, (size_t)cfg->n_nodes * sizeof(uint32_t)));
        CK(cudaMemset(c->node_hotness, 0, (size_t)cfg->n_nodes * sizeof(uint32_t)));
        CK(cudaMemset(c->topo_hotness, 0, (size_t)cfg->n_nodes * sizeof(uint32_t)));
    }
    {
        const char* gm = getenv("LGN_GATHER");
        c->gather_mode = !gm ? -1 : (gm[0] == 'l' ? 0 : 1);   // -1 = auto (by tier mix)
        const char* gc = getenv("LGN_GATHER_CTAS");
        c->gather_ctas_per_sm = gc ? atoi(gc) : 1;
        if (c->gather_ctas_per_sm < 1) c->gather_ctas_per_sm = 1;
        auto knob = [](const char* name, int dflt) { const char* v = getenv(name); int x = v ? atoi(v) : dflt; return x < 1 ? 1 : x; };
        c->gather_ldg_ctas = getenv("LGN_GATHER_LDG_CTAS") ? knob("LGN_GATHER_LDG_CTAS", 8) : 0;   // 0 = auto
        c->gather_unroll = knob("LGN_GATHER_UNROLL", 4);
        c->gather_threads = knob("LGN_GATHER_THREADS", 256);
        if (c->gather_threads > 256) c->gather_threads = 256;
        c->gather_threads = (c->gather_threads + 31) / 32 * 32;
        c->sample_ctas_per_sm = knob("LGN_SAMPLE_CTAS", 16);
        c->resolve_ctas_per_sm = knob("LGN_RESOLVE_CTAS", 12);   // 128-thread CTAs, one tile per iteration
        c->end_ctas_per_sm = knob("LGN_END_CTAS", 4);
        const char* ug = getenv("LGN_GRAPH");
        c->use_graphs = ug ? atoi(ug) : 1;
        const char* lp = getenv("LGN_L2_PERSIST");
        c->l2_persist = lp ? atoi(lp) : 0;
        if (c->l2_persist) {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, cfg->device));
            size_t want = (size_t)c->n_lanes * (c->dedup_hash ? ((size_t)8 << c->dedup_bits_max) : (size_t)cfg->n_nodes * 4);
            if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        }
        if (const char* fg = getenv("LGN_L2_FETCH")) {      // experiment: L2 fetch granularity hint (32 / 64 / 128 bytes)
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg));
            cudaGetLastError();
        }
        const char* sg = getenv("LGN_SHARED_GATHER");
        c->shared_gather_stream = sg ? atoi(sg) : 0;
    }
    c->feat.my_part = cfg->part;
    lgn::gather_init_device(c);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return LGN_OK;
}

int lgn_destroy(lgn_ctx* c)
{
    if (!c) return LGN_E_ARG;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    for (int p = 0; p < c->n_lanes; p++) {
        lgn::Pipe& pp = c->pipe[p];
        void* wire[7] = {pp.ids, pp.features, pp.labels, pp.agg_src_off, pp.agg_dst_off, pp.nc, pp.ec};
        for (int i = 0; i < 7; i++) if (!(pp.external & (1u << i))) cudaFree(wire[i]);     // attached buffers belong to the caller
        cudaFree(pp.slot_map); cudaFree(pp.dedup_tab); cudaFree(pp.seed_h); cudaFree(pp.agg_src_ids); cudaFree(pp.agg_dst_ids);
        cudaFree(pp.draw_h); cudaFree(pp.draw_s); cudaFree(pp.draw_v); cudaFree(pp.draw_key); cudaFree(pp.tile_n); cudaFree(pp.tile_new);
        cudaFree(pp.super_e); cudaFree(pp.super_n); cudaFree(pp.state); cudaFree(pp.seed_stage);
        if (pp.gather_stream) cudaStreamDestroy(pp.gather_stream);
        for (int i = 0; i < LGN_MAX_HOPS + 2; i++) if (pp.ev_hop[i]) cudaEventDestroy(pp.ev_hop[i]);
        if (pp.ev_end) cudaEventDestroy(pp.ev_end);
        if (pp.ev_done) cudaEventDestroy(pp.ev_done);
        if (pp.ev_join) cudaEventDestroy(pp.ev_join);
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) if (pp.graph_exec[a][b]) cudaGraphExecDestroy(pp.graph_exec[a][b]);
    }
    cudaFree(c->node_hotness); cudaFree(c->topo_hotness);
    lgn_profile_enable(c, 0);
    cudaGetLastError();
    delete c;
    return LGN_OK;
}

int64_t lgn_capacity(const lgn_ctx* c) { return c ? c->capacity : 0; }

int lgn_set_dedup_capacity(lgn_ctx* c, int64_t expected_unique)
{
    if (!c || expected_unique <= 0) return LGN_E_ARG;
    if (!c->dedup_hash) return LGN_OK;
    for (int i = 0; i < c->n_lanes; i++) if (c->pipe[i].pending) CK(cudaEventSynchronize(c->pipe[i].ev_done));
    uint32_t bits = 10;
    while (((long long)1 << bits) < (5 * expected_unique) / 2 && bits < c->dedup_bits_max) bits++;
    for (int i = 0; i < c->n_lanes; i++) {      // entries move when the table shrinks or grows: start from an empty one
        c->pipe[i].dedup.bits = bits;
        CK(cudaMemset(c->pipe[i].dedup_tab, 0xFF, sizeof(unsigned long long) << bits));
    }
    invalidate_graphs(c);
    return LGN_OK;
}

int lgn_set_epoch(lgn_ctx* c, uint32_t epoch, uint32_t step_offset)
{
    if (!c) return LGN_E_ARG;
    c->rng_epoch = epoch;               // kernel arguments of the next k_batch_begin: no captured graph depends on them
    c->rng_step_offset = step_offset;
    return LGN_OK;
}

int lgn_set_part(lgn_ctx* c, int32_t part)
{
    if (!c || part < 0 || part >= LGN_MAX_PARTS) return LGN_E_ARG;
    c->cfg.part = part;
    c->feat.my_part = part;
    invalidate_graphs(c);
    return LGN_OK;
}

// ---------------------------------------------------------------- storage binding
static void invalidate_graphs_impl(lgn_ctx* c)
{
    for (int p = 0; p < c->n_lanes; p++)
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                if (c->pipe[p].graph_exec[a][b]) { cudaGraphExecDestroy(c->pipe[p].graph_exec[a][b]); c->pipe[p].graph_exec[a][b] = nullptr; }
                c->pipe[p].graph_calls[a][b] = 0;
            }
}

int lgn_bind_seeds(lgn_ctx* c, int32_t mode, const int32_t* ids, const int32_t* labels, int32_t count)
{
    if (!c || mode < 0 || mode > 2 || count < 0 || (count > 0 && !ids)) return LGN_E_ARG;
    c->seed_ids[mode] = ids; c->seed_labels[mode] = labels; c->seed_count[mode] = count;
    return LGN_OK;
}
int lgn_bind_topology(lgn_ctx* c, const int64_t* indptr, const int32_t* indices)
{
    if (!c || !indptr || !indices) return LGN_E_ARG;
    c->topo.base_indptr = indptr; c->topo.base_indices = indices;
    invalidate_graphs(c);
    return LGN_OK;
}
int lgn_bind_topology_cache(lgn_ctx* c, int32_t n_parts, const int64_t* const* indptr_tab, const int32_t* const* indices_tab,
                            const int32_t* slot_of, int64_t cap)
{
    if (!c || n_parts < 0 || n_parts > LGN_MAX_PARTS) return LGN_E_ARG;
    invalidate_graphs(c);
    if (!slot_of || n_parts == 0) { c->topo.slot_of = nullptr; return LGN_OK; }
    if (cap <= 0 || !indptr_tab || !indices_tab) return LGN_E_ARG;
    for (int i = 0; i < n_parts; i++) { c->topo.indptr_tab[i] = indptr_tab[i]; c->topo.indices_tab[i] = indices_tab[i]; }
    c->topo.cap = cap; c->topo.slot_of = slot_of;
    return LGN_OK;
}
int lgn_bind_features(lgn_ctx* c, const float* features)
{
    if (!c || !features) return LGN_E_ARG;
    c->feat.base = features;
    invalidate_graphs(c);
    return LGN_OK;
}
int lgn_bind_feature_cache(lgn_ctx* c, int32_t n_parts, const float* const* shard_tab, const int32_t* slot_of, int64_t cap)
{
    if (!c || n_parts < 0 || n_parts > LGN_MAX_PARTS) return LGN_E_ARG;
    invalidate_graphs(c);
    c->feat.cmap = nullptr; c->feat.identity = 0;
    if (!slot_of || n_parts == 0) { c->feat.slot_of = nullptr; c->feat.n_parts = 0; return LGN_OK; }
    if (cap <= 0 || !shard_tab) return LGN_E_ARG;
    for (int i = 0; i < n_parts; i++) c->feat.shard_tab[i] = shard_tab[i];
    c->feat.cap = cap; c->feat.slot_of = slot_of; c->feat.n_parts = n_parts;
    return LGN_OK;
}
int lgn_bind_feature_cache_compact(lgn_ctx* c, int32_t n_parts, const float* const* shard_tab, const void* cmap, int64_t n_repl, int64_t cap)
{
    if (!c || n_parts <= 0 || n_parts > LGN_MAX_PARTS || !shard_tab || n_repl < 0 || cap <= 0 || n_repl > cap) return LGN_E_ARG;
    if (c->feat.my_part >= n_parts) return LGN_E_ARG;
    if (!cmap && !(n_repl == c->cfg.n_nodes && cap >= n_repl)) return LGN_E_ARG;   // direct addressing needs the whole table in shard_tab[my_part]
    if (cmap && ((uintptr_t)cmap & 15)) return LGN_E_ARG;
    invalidate_graphs(c);
    for (int i = 0; i < n_parts; i++) { if (!shard_tab[i]) return LGN_E_ARG; c->feat.shard_tab[i] = shard_tab[i]; }
    c->feat.slot_of = nullptr;
    c->feat.cmap = (const uint4*)cmap;
    c->feat.identity = cmap ? 0 : 1;
    c->feat.n_repl = n_repl; c->feat.kg = n_parts; c->feat.cap = cap; c->feat.n_parts = n_parts;
    return LGN_OK;
}

// ---------------------------------------------------------------- L2 persistence
// The lane's dedup structure is touched at random several times per sampled edge while the gathers stream
// hundreds of MB through L2: give it a persisting access-policy window on the lane's sampling stream
// (hardware-managed set-aside, cudaLimitPersistingL2CacheSize) on top of the per-access evict_last hints.
static int apply_l2_window(lgn_ctx* c, int pipe, cudaStream_t s)
{
    lgn::Pipe& pp = c->pipe[pipe];
    if (!c->l2_persist || !s || pp.window_stream == s) return LGN_OK;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    size_t bytes = c->dedup_hash ? ((size_t)8 << pp.dedup.bits) : (size_t)c->cfg.n_nodes * 4;
    int dev = c->cfg.device, max_win = 0;
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;
    v.accessPolicyWindow.base_ptr = c->dedup_hash ? (void*)pp.dedup_tab : (void*)pp.slot_map;
    v.accessPolicyWindow.num_bytes = bytes;
    v.accessPolicyWindow.hitRatio = 1.0f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
    pp.window_stream = s;
    return LGN_OK;
}

// ---------------------------------------------------------------- operator timing
struct ProfScope {
    lgn_ctx* c; cudaStream_t s; int idx;
    ProfScope(lgn_ctx* c_, cudaStream_t s_, int kind) : c(c_), s(s_), idx(-1)
    {
        if (c->prof_cap && c->prof_n < c->prof_cap) {
            idx = c->prof_n++;
            c->prof_kind[idx] = (signed char)kind;
            c->prof_pipe[idx] = (signed char)c->cur_pipe;
            cudaEventRecord(c->prof_ev[2 * idx], s);
        }
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof_ev[2 * idx + 1], s); }
};

int lgn_profile_enable(lgn_ctx* c, int32_t max_records)
{
    if (!c || max_records < 0) return LGN_E_ARG;
    for (int i = 0; i < 2 * c->prof_cap; i++) cudaEventDestroy(c->prof_ev[i]);
    delete[] c->prof_ev; delete[] c->prof_kind; delete[] c->prof_pipe;
    c->prof_ev = nullptr; c->prof_kind = nullptr; c->prof_pipe = nullptr; c->prof_cap = c->prof_n = 0;
    if (max_records == 0) return LGN_OK;
    c->prof_ev = new cudaEvent_t[2 * (size_t)max_records];
    c->prof_kind = new signed char[max_records];
    c->prof_pipe = new signed char[max_records];
    for (int i = 0; i < 2 * max_records; i++) {
        cudaError_t e = cudaEventCreate(&c->prof_ev[i]);
        if (e != cudaSuccess) {          // release what exists: a failed enable leaves profiling off, nothing leaked
            for (int k = 0; k < i; k++) cudaEventDestroy(c->prof_ev[k]);
            delete[] c->prof_ev; delete[] c->prof_kind; delete[] c->prof_pipe;
            c->prof_ev = nullptr; c->prof_kind = nullptr; c->prof_pipe = nullptr;
            return lgn_cuda_fail(e, "cudaEventCreate");
        }
    }
    c->prof_cap = max_records;
    return LGN_OK;
}

int lgn_profile_collect(lgn_ctx* c, double ms[4], int64_t calls[4])
{
    if (!c || !ms || !calls) return LGN_E_ARG;
    for (int k = 0; k < 4; k++) { ms[k] = 0.0; calls[k] = 0; }
    for (int i = 0; i < c->prof_n; i++) {
        CK(cudaEventSynchronize(c->prof_ev[2 * i + 1]));
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
        ms[c->prof_kind[i] & 3] += t;
        calls[c->prof_kind[i] & 3]++;
    }
    c->prof_n = 0;
    return LGN_OK;
}

int lgn_profile_timeline(lgn_ctx* c, double* rows, int32_t max_records, int32_t* n_out)
{
    if (!c || !rows || !n_out) return LGN_E_ARG;
    int n = c->prof_n < max_records ? c->prof_n : max_records;
    for (int i = 0; i < n; i++) {
        CK(cudaEventSynchronize(c->prof_ev[2 * i + 1]));
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, c->prof_ev[0], c->prof_ev[2 * i]));
        CK(cudaEventElapsedTime(&b, c->prof_ev[0], c->prof_ev[2 * i + 1]));
        rows[4 * i] = c->prof_kind[i]; rows[4 * i + 1] = c->prof_pipe[i]; rows[4 * i + 2] = a; rows[4 * i + 3] = b;
    }
    *n_out = n;
    return LGN_OK;
}

int lgn_debug_shard_read(lgn_ctx* c, void* stream, int32_t slot, int64_t n_rows, int64_t rows_per_shard, int32_t peers_only,
                         int32_t repeats, double* avg_ms)
{
    if (!c || !avg_ms || slot < 0 || slot >= c->n_lanes || n_rows <= 0 || rows_per_shard <= 0 || repeats <= 0) return LGN_E_ARG;
    if (!c->feat.slot_of || c->feat.n_parts == 0 || c->cfg.feat_dim <= 0 || (c->cfg.feat_dim & 3) || c->cfg.feat_dim > 512) return LGN_E_STATE;
    if (n_rows > c->max_rows || rows_per_shard > c->feat.cap || rows_per_shard > 0xffffffffLL) return LGN_E_CAPACITY;
    if (peers_only && c->feat.n_parts < 2) return LGN_E_STATE;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch_debug_shard_read(c, s, slot, n_rows, rows_per_shard, peers_only != 0, 1u);       // warm-up
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < repeats; i++) launch_debug_shard_read(c, s, slot, n_rows, rows_per_shard, peers_only != 0, 0x9e3779b9u * (uint32_t)(i + 2));
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    *avg_ms = (double)ms / repeats;
    return LGN_OK;
}

// ---------------------------------------------------------------- hot path
int lgn_batch_generate(lgn_ctx* c, void* stream, int32_t pipe, int32_t mode, int32_t batch_size, int32_t counter)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes || mode < 0 || mode > 2 || batch_size < 0 || counter < 0) return LGN_E_ARG;   // 0 = empty batch (a partition without valid/test ids)
    if (batch_size > c->cfg.batch_size) return LGN_E_CAPACITY;
    if (!c->seed_ids[mode]) return LGN_E_STATE;
    const int32_t total = c->seed_count[mode];
    // Kernels.cu:224-227: last batch is clamped, and the kernel strides by the clamped size
    long long size = ((long long)batch_size * (counter + 1) >= total) ? (long long)total - (long long)batch_size * counter : batch_size;
    if (size < 0) size = 0;
    const long long off = size * counter;
    c->cur_pipe = pipe;
    if (c->pipe[pipe].pending) CK(cudaStreamWaitEvent((cudaStream_t)stream, c->pipe[pipe].ev_done, 0));   // slot reuse (WAR)
    { int rcw = apply_l2_window(c, pipe, (cudaStream_t)stream); if (rcw) return rcw; }
    ProfScope prof(c, (cudaStream_t)stream, 0);
    launch_batch_begin(c, (cudaStream_t)stream, c->seed_ids[mode], c->seed_labels[mode], (int32_t)off, (int32_t)size, (uint32_t)counter);
    CK(cudaGetLastError());
    return LGN_OK;
}

int lgn_batch_from_host(lgn_ctx* c, void* stream, int32_t pipe, const int32_t* seeds, const int32_t* labels, int32_t count, uint32_t step)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes || count < 0 || (count > 0 && !seeds)) return LGN_E_ARG;
    if (count > c->cfg.batch_size) return LGN_E_CAPACITY;
    cudaStream_t s = (cudaStream_t)stream;
    c->cur_pipe = pipe;
    if (c->pipe[pipe].pending) CK(cudaStreamWaitEvent(s, c->pipe[pipe].ev_done, 0));   // slot reuse (WAR)
    { int rcw = apply_l2_window(c, pipe, s); if (rcw) return rcw; }
    int32_t* stage = c->pipe[pipe].seed_stage;
    if (count > 0) CK(cudaMemcpyAsync(stage, seeds, (size_t)count * 4, cudaMemcpyHostToDevice, s));
    if (count > 0 && labels) CK(cudaMemcpyAsync(stage + c->cfg.batch_size, labels, (size_t)count * 4, cudaMemcpyHostToDevice, s));
    ProfScope prof(c, s, 0);
    launch_batch_begin(c, s, stage, labels ? stage + c->cfg.batch_size : nullptr, 0, count, step);
    CK(cudaGetLastError());
    return LGN_OK;
}

int lgn_sample_hop(lgn_ctx* c, void* stream, int32_t hop, int32_t is_presc)
{
    if (!c || hop < 0 || hop >= c->cfg.n_hops) return LGN_E_ARG;
    if (!c->topo.base_indptr) return LGN_E_STATE;
    if (is_presc && !c->topo_hotness) return LGN_E_STATE;
    ProfScope prof(c, (cudaStream_t)stream, 1);
    launch_sample_hop(c, (cudaStream_t)stream, hop, is_presc != 0);
    CK(cudaGetLastError());
    return LGN_OK;
}

int lgn_gather_segment(lgn_ctx* c, void* stream, int32_t segment)
{
    if (!c || segment < 0 || segment > c->cfg.n_hops) return LGN_E_ARG;
    if (!c->feat.base || c->cfg.feat_dim <= 0) return LGN_E_STATE;
    ProfScope prof(c, (cudaStream_t)stream, 2);
    launch_gather(c, (cudaStream_t)stream, segment, 1);
    CK(cudaGetLastError());
    return LGN_OK;
}

int lgn_gather_segments(lgn_ctx* c, void* stream, int32_t first, int32_t n)
{
    if (!c || first < 0 || n < 1 || first + n > c->cfg.n_hops + 1) return LGN_E_ARG;
    if (!c->feat.base || c->cfg.feat_dim <= 0) return LGN_E_STATE;
    ProfScope prof(c, (cudaStream_t)stream, 2);
    launch_gather(c, (cudaStream_t)stream, first, n);
    CK(cudaGetLastError());
    return LGN_OK;
}

// the feature-extraction launches of one batch, as lgn_run_batch issues them: the seeds are fused with hop 1's new
// nodes (two small, latency-bound gathers become one), then one launch per further hop
int lgn_gather_batch(lgn_ctx* c, void* stream)
{
    if (!c) return LGN_E_ARG;
    if (!c->feat.base || c->cfg.feat_dim <= 0) return LGN_E_STATE;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->cfg.n_hops == 0) return lgn_gather_segment(c, stream, 0);
    for (int h = 0; h < c->cfg.n_hops; h++) {
        ProfScope prof(c, s, 2);
        if (h == 0) launch_gather(c, s, 0, 2); else launch_gather(c, s, h + 1, 1);
        CK(cudaGetLastError());
    }
    return LGN_OK;
}

int32_t lgn_launches_per_batch(const lgn_ctx* c, int32_t with_features)
{
    if (!c) return 0;
    const int hops = c->cfg.n_hops;
    // k_batch_begin + (k_sample, k_mark, k_assign) per hop + k_batch_end, + the gathers of lgn_gather_batch
    return 2 + 3 * hops + (with_features ? (hops == 0 ? 1 : hops) : 0);
}

const char* lgn_gather_kernel_name(const lgn_ctx* c)
{
    if (!c) return "";
    return gather_kernel_name(c);
}

int lgn_finish_batch(lgn_ctx* c, void* stream, int32_t is_presc)
{
    if (!c) return LGN_E_ARG;
    if (is_presc && !c->node_hotness) return LGN_E_STATE;
    ProfScope prof(c, (cudaStream_t)stream, 3);
    launch_batch_end(c, (cudaStream_t)stream, is_presc != 0);
    CK(cudaGetLastError());
    return LGN_OK;
}

// GPURunner::RunOnce / RunPreSc (Server.cu:284-328): sampling ops on `stream`, feature
// extraction ops on the second stream, chained by events.  Unlike the reference (which
// busy-polls the last event before touching the next batch, Server.cu:318-323) the caller's
// stream is NOT joined here: the next batch's sampling overlaps this batch's gathers, the
// slot's completion is the event lgn_wait_pipe / lgn_read_counters wait on.
static int enqueue_batch(lgn_ctx* c, cudaStream_t s, bool feats, int32_t is_presc, bool join)
{
    lgn::Pipe& pp = c->pipe[c->cur_pipe];
    cudaStream_t g = c->shared_gather_stream ? c->pipe[0].gather_stream : pp.gather_stream;
    int rc;
    // feature extraction of the seeds is fused with hop 1's segment (two small, latency-bound gathers
    // become one); with no hops at all the seeds are gathered alone
    if (feats && c->cfg.n_hops == 0) {
        CK(cudaEventRecord(pp.ev_hop[0], s));
        CK(cudaStreamWaitEvent(g, pp.ev_hop[0], 0));
        if ((rc = lgn_gather_segment(c, g, 0))) return rc;
    }
    for (int h = 0; h < c->cfg.n_hops; h++) {
        if ((rc = lgn_sample_hop(c, s, h, is_presc))) return rc;
        if (feats) {
            CK(cudaEventRecord(pp.ev_hop[h + 1], s));
            CK(cudaStreamWaitEvent(g, pp.ev_hop[h + 1], 0));
            ProfScope prof(c, g, 2);
            if (h == 0) launch_gather(c, g, 0, 2); else launch_gather(c, g, h + 1, 1);
            CK(cudaGetLastError());
        }
    }
    if ((rc = lgn_finish_batch(c, s, is_presc))) return rc;
    if (join) {          // captured form: the gather branch must rejoin the origin stream
        if (feats) {
            CK(cudaEventRecord(pp.ev_join, g));
            CK(cudaStreamWaitEvent(s, pp.ev_join, 0));
        }
    } else if (feats) {   // slot complete = last gather done AND batch end done
        CK(cudaEventRecord(pp.ev_end, s));
        CK(cudaStreamWaitEvent(g, pp.ev_end, 0));
        CK(cudaEventRecord(pp.ev_done, g));
    } else {
        CK(cudaEventRecord(pp.ev_done, s));
    }
    return LGN_OK;
}

int lgn_run_batch(lgn_ctx* c, void* stream, int32_t with_features, int32_t is_presc)
{
    if (!c) return LGN_E_ARG;
    lgn::Pipe& pp = c->pipe[c->cur_pipe];
    cudaStream_t s = (cudaStream_t)stream;
    const bool feats = with_features && !is_presc && c->feat.base && c->cfg.feat_dim > 0;
    if (with_features && !is_presc && !feats) return LGN_E_STATE;
    const int fi = feats ? 1 : 0, pi = is_presc ? 1 : 0;
    // CUDA graph replay: the per-batch DAG has no host-visible parameters (every count lives in device memory), so it
    // is captured once per slot and replayed.  Needs a real stream (the legacy default stream cannot be captured),
    // no operator timing, and one eager call first (lazy module loading / attribute setup are not capturable).
    const bool graphable = c->use_graphs && s != nullptr && c->prof_cap == 0 && !c->shared_gather_stream;
    if (graphable && pp.graph_calls[fi][pi] >= 1) {
        if (!pp.graph_exec[fi][pi]) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_batch(c, s, feats, is_presc, true);
            cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return lgn_cuda_fail(e, "cudaStreamEndCapture");
            e = cudaGraphInstantiate(&pp.graph_exec[fi][pi], graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return lgn_cuda_fail(e, "cudaGraphInstantiate");
        }
        CK(cudaGraphLaunch(pp.graph_exec[fi][pi], s));
        CK(cudaEventRecord(pp.ev_done, s));
    } else {
        int rc = enqueue_batch(c, s, feats, is_presc, false);
        if (rc) return rc;
        pp.graph_calls[fi][pi]++;
    }
    pp.pending = true;
    return LGN_OK;
}

int lgn_select_pipe(lgn_ctx* c, int32_t pipe)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    c->cur_pipe = pipe;
    return LGN_OK;
}

int lgn_wait_pipe(lgn_ctx* c, void* stream, int32_t pipe)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    if (c->pipe[pipe].pending) CK(cudaStreamWaitEvent((cudaStream_t)stream, c->pipe[pipe].ev_done, 0));
    return LGN_OK;
}

int lgn_sync_pipe(lgn_ctx* c, int32_t pipe)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    if (c->pipe[pipe].pending) CK(cudaEventSynchronize(c->pipe[pipe].ev_done));
    return LGN_OK;
}

// ---------------------------------------------------------------- results
int lgn_batch_buffers(lgn_ctx* c, int32_t pipe, lgn_batch_view* v)
{
    if (!c || !v || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    const lgn::Pipe& p = c->pipe[pipe];
    v->ids = p.ids; v->features = p.features; v->labels = p.labels; v->agg_src = p.agg_src_off; v->agg_dst = p.agg_dst_off;
    v->node_counter = p.nc; v->edge_counter = p.ec; v->agg_src_ids = p.agg_src_ids; v->agg_dst_ids = p.agg_dst_ids;
    v->capacity = c->capacity; v->max_rows = c->max_rows;
    return LGN_OK;
}

// A reference-style runner owns the seven wire buffers of a slot itself (Server.cu:217-283 allocates them, or takes
// them from IPCEnv, and registers them in its GPUMemoryPool).  Every non-NULL pointer of *v replaces the slot's own
// buffer (which is released); the caller keeps ownership and must keep them alive until lgn_destroy.  Sizes: ids and
// the two edge arrays hold v->capacity entries, features v->max_rows rows; smaller buffers than the context's worst
// case lower its limits (the reference sizes them from the raw batch and 1.2 x max_ids): overflow sets LGN_E_CAPACITY.
int lgn_attach_buffers(lgn_ctx* c, int32_t pipe, const lgn_batch_view* v)
{
    if (!c || !v || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    if ((v->ids || v->agg_src || v->agg_dst) && v->capacity < c->cfg.batch_size) return LGN_E_CAPACITY;   // not even the seeds fit
    if (v->features && v->max_rows <= 0) return LGN_E_ARG;
    if (((uintptr_t)v->agg_src | (uintptr_t)v->features) & 15) return LGN_E_ARG;      // 128-bit accesses (k_batch_end, the gathers)
    lgn::Pipe& pp = c->pipe[pipe];
    if (pp.pending) CK(cudaEventSynchronize(pp.ev_done));
    auto swap = [&](int bit, void** slot, void* ext) {
        if (!ext || *slot == ext) return;
        if (!(pp.external & (1u << bit))) cudaFree(*slot);
        *slot = ext;
        pp.external |= 1u << bit;
    };
    swap(0, (void**)&pp.ids, v->ids);
    swap(1, (void**)&pp.features, v->features);
    swap(2, (void**)&pp.labels, v->labels);
    swap(3, (void**)&pp.agg_src_off, v->agg_src);
    swap(4, (void**)&pp.agg_dst_off, v->agg_dst);
    swap(5, (void**)&pp.nc, v->node_counter);
    swap(6, (void**)&pp.ec, v->edge_counter);
    if (v->features && v->max_rows < c->max_rows) c->max_rows = v->max_rows;
    if ((v->ids || v->agg_src || v->agg_dst) && v->capacity < c->capacity) c->capacity = v->capacity;   // kernels bound every id / edge write by it
    invalidate_graphs(c);
    return LGN_OK;
}

int lgn_read_counters(lgn_ctx* c, void* stream, int32_t pipe, int32_t nc[16], int32_t ec[16])
{
    if (!c || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->pipe[pipe].pending) CK(cudaStreamWaitEvent(s, c->pipe[pipe].ev_done, 0));
    CK(cudaMemcpyAsync(nc, c->pipe[pipe].nc, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ec, c->pipe[pipe].ec, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return LGN_OK;
}

static int sync_lanes(lgn_ctx* c)
{
    for (int i = 0; i < c->n_lanes; i++) {
        if (c->pipe[i].pending) CK(cudaEventSynchronize(c->pipe[i].ev_done));
        CK(cudaStreamSynchronize(c->pipe[i].gather_stream));
    }
    return LGN_OK;
}

int lgn_tier_counts(lgn_ctx* c, void* stream, int64_t out[3], int32_t reset)
{
    if (!c || !out) return LGN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = sync_lanes(c);
    if (rc) return rc;
    out[0] = out[1] = out[2] = 0;
    for (int i = 0; i < c->n_lanes; i++) {
        unsigned long long h[4];
        CK(cudaMemcpyAsync(h, c->pipe[i].state->tier_rows, sizeof(h), cudaMemcpyDeviceToHost, s));
        if (reset) CK(cudaMemsetAsync(c->pipe[i].state->tier_rows, 0, sizeof(h), s));
        CK(cudaStreamSynchronize(s));
        out[0] += (int64_t)h[0]; out[1] += (int64_t)h[1]; out[2] += (int64_t)h[2];
    }
    return LGN_OK;
}

// host-blocking: waits for the slot's batch and returns ITS sticky status (0 or LGN_E_CAPACITY): what a server checks
// before it hands the slot to a trainer
int lgn_sync_pipe_status(lgn_ctx* c, int32_t pipe)
{
    if (!c || pipe < 0 || pipe >= c->n_lanes) return LGN_E_ARG;
    if (c->pipe[pipe].pending) CK(cudaEventSynchronize(c->pipe[pipe].ev_done));
    int32_t st = 0;
    CK(cudaMemcpy(&st, &c->pipe[pipe].state->status, 4, cudaMemcpyDeviceToHost));
    return st;
}

int lgn_status(lgn_ctx* c, void* stream)
{
    if (!c) return LGN_E_ARG;
    int rc = sync_lanes(c);
    if (rc) return rc;
    int32_t worst = 0;
    for (int i = 0; i < c->n_lanes; i++) {
        int32_t st = 0;
        CK(cudaMemcpyAsync(&st, &c->pipe[i].state->status, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CK(cudaStreamSynchronize((cudaStream_t)stream));
        if (st) worst = st;
    }
    return worst;
}

int lgn_sampling_totals(lgn_ctx* c, void* stream, int64_t out[2], int32_t reset)
{
    if (!c || !out) return LGN_E_ARG;
    int rc = sync_lanes(c);
    if (rc) return rc;
    out[0] = out[1] = 0;
    for (int i = 0; i < c->n_lanes; i++) {
        unsigned long long h[2];
        CK(cudaMemcpyAsync(h, &c->pipe[i].state->tot_items, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        if (reset) CK(cudaMemsetAsync(&c->pipe[i].state->tot_items, 0, sizeof(h), (cudaStream_t)stream));
        CK(cudaStreamSynchronize((cudaStream_t)stream));
        out[0] += (int64_t)h[0]; out[1] += (int64_t)h[1];
    }
    return LGN_OK;
}

__global__ void k_accumulate_u32(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

int lgn_accumulate_u32(uint32_t* dst, const uint32_t* src, int64_t n, void* stream)
{
    if (!dst || !src || n < 0) return LGN_E_ARG;
    k_accumulate_u32<<<1184, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    CK(cudaGetLastError());
    return LGN_OK;
}

int lgn_hotness(lgn_ctx* c, uint32_t** node, uint32_t** topo)
{
    if (!c || !c->node_hotness) return LGN_E_STATE;
    if (node) *node = c->node_hotness;
    if (topo) *topo = c->topo_hotness;
    return LGN_OK;
}

int32_t lgn_max_ids(lgn_ctx* c, void* stream)
{
    if (!c) return LGN_E_ARG;
    if (sync_lanes(c)) return LGN_E_CUDA;
    int32_t best = 0;
    for (int i = 0; i < c->n_lanes; i++) {
        int32_t v = 0;
        if (cudaMemcpyAsync(&v, &c->pipe[i].state->max_ids, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return LGN_E_CUDA;
        cudaStreamSynchronize((cudaStream_t)stream);
        if (v > best) best = v;
    }
    return best;
}

// ---------------------------------------------------------------- step arithmetic
int lgn_coordinate(const int32_t* n_train, const int32_t* n_valid, const int32_t* n_test, int32_t parts, int32_t batch,
                   int32_t epochs, lgn_steps* out)
{
    if (!n_train || !n_valid || !n_test || !out || parts <= 0 || parts > LGN_MAX_PARTS || batch <= 0) return LGN_E_ARG;
    int32_t min_train = 1000000000, max_valid = 0, max_test = 0;      // CUDA_IPC_Service.cu:71-112
    for (int i = 0; i < parts; i++) {
        if (n_train[i] < min_train) min_train = n_train[i];
        if (n_valid[i] > max_valid) max_valid = n_valid[i];
        if (n_test[i] > max_test) max_test = n_test[i];
    }
    memset(out, 0, sizeof(*out));
    out->train_step = (min_train - 1) / batch;
    out->valid_step = (max_valid - 1) / 512 + 1;
    out->test_step = (max_test - 1) / 512 + 1;
    for (int i = 0; i < parts; i++) {
        out->valid_batch[i] = (n_valid[i] - 1) / out->valid_step + 1;
        out->test_batch[i] = (n_test[i] - 1) / out->test_step + 1;
    }
    out->max_step = (out->train_step + out->valid_step) * epochs + out->test_step;   // :136-138
    return LGN_OK;
}

int32_t lgn_mode_of_step(const lgn_steps* s, int32_t epochs, int32_t g)
{
    if (g < (s->train_step + s->valid_step) * epochs)                                 // :246-259
        return (g % (s->train_step + s->valid_step)) < s->train_step ? LGN_MODE_TRAIN : LGN_MODE_VALID;
    return LGN_MODE_TEST;
}

int32_t lgn_local_batch_id(const lgn_steps* s, int32_t epochs, int32_t g)
{
    if (g < (s->train_step + s->valid_step) * epochs) {                               // :219-233
        const int32_t e = g % (s->train_step + s->valid_step);
        return e < s->train_step ? e : e - s->train_step;
    }
    return (g - (s->train_step + s->valid_step) * epochs) % s->test_step;
}

}  // extern "C"

// start of a generation cycle: every value of the lane's dedup structure back to EMPTY (every 63rd batch of a slot)
namespace lgn {
void reset_dedup(lgn_ctx* c, Pipe& p, cudaStream_t s)
{
    if (c->dedup_hash) cudaMemsetAsync(p.dedup_tab, 0xFF, sizeof(unsigned long long) << p.dedup.bits, s);
    else k_fill_i32<<<1024, 256, 0, s>>>(p.slot_map, lgn::EMPTY, c->cfg.n_nodes);
}
}  // namespace lgn

