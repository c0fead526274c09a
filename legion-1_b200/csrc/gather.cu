// gather.cu -- feature extraction from the tiered cache for sm_100a.
//
// Replaces zero_copy_with_aggregated_cache (Kernels.cu:662-702), the cuckoo lookup of
// FindFeat (GPUCache.cu:387-432, bght::bcht::find) and FeatFillUp (GPUCache.cu:200-205).
//
// One warp owns 32 consecutive output rows: lane l resolves row l's tier (resolve_row: no
// lookup at all for a resident table, a 32-byte compact-map record, or a direct-mapped int32
// slot table -- instead of up to three 128-byte cuckoo buckets), the
// warp then streams the rows UNROLL at a time with 128-bit loads -- local HBM shard,
// peer shard over NVLink (P2P load) or mapped host memory over PCIe (UVA zero-copy) are
// all plain global addresses -- and writes them with 128-bit streaming stores.  The
// reference runs one thread per float with a 64-bit divide and modulo per element.
#include <stdlib.h>

#include "context.h"

namespace lgn {

constexpr int GATHER_THREADS = 256;
constexpr int GATHER_UNROLL = 4;

// ---- row -> tier + address --------------------------------------------------------------
// Three bindings (context.h: FeatView):
//   identity  the whole table is resident in id order: no lookup at all.
//   compact   lgn_place_compact: a 32-byte record per 96 nodes {repl_before, part_before, repl_bits[3], part_bits[3]}.  The
//             map of a 111 M-node graph is 37 MB and stays in L2 (evict_last), where the int32 slot table (444 MB) costs one
//             random DRAM access per gathered row -- measured 9 % of the whole step on the papers100M shape (DESIGN.md
//             section 4).  Rows of a class are stored in node-id order: row = prefix + popcount.
//   slot_of   int32[N] part*cap+row / -1 (the reference placement, hotness order inside the shards).
// tier: 0 local shard, 1 peer shard, 2 base matrix (host).
__device__ __forceinline__ const float* resolve_row(const FeatView& fv, long long nid, int dim, int& tier, unsigned long long keep)
{
    if (fv.identity) { tier = 0; return fv.shard_tab[fv.my_part] + nid * dim; }
    if (fv.cmap) {
        const uint32_t rec = (uint32_t)nid / (uint32_t)LGN_CMAP_NODES, j = (uint32_t)nid - rec * (uint32_t)LGN_CMAP_NODES;
        const uint4 a = ld_stream_v4(fv.cmap + 2 * (size_t)rec, keep), b = ld_stream_v4(fv.cmap + 2 * (size_t)rec + 1, keep);
        const uint32_t wi = j >> 5, bit = j & 31u, below = (1u << bit) - 1u;
        const uint32_t rw = wi == 0 ? a.z : (wi == 1 ? a.w : b.x);
        const uint32_t pw = wi == 0 ? b.y : (wi == 1 ? b.z : b.w);
        if ((rw >> bit) & 1u) {
            const uint32_t row = a.x + (wi > 0 ? __popc(a.z) : 0) + (wi > 1 ? __popc(a.w) : 0) + __popc(rw & below);
            tier = 0;
            return fv.shard_tab[fv.my_part] + (long long)row * dim;
        }
        if ((pw >> bit) & 1u) {
            const uint32_t q = a.y + (wi > 0 ? __popc(b.y) : 0) + (wi > 1 ? __popc(b.z) : 0) + __popc(pw & below);
            const uint32_t part = q % (uint32_t)fv.kg;
            tier = (int)part == fv.my_part ? 0 : 1;
            return fv.shard_tab[part] + (fv.n_repl + (long long)(q / (uint32_t)fv.kg)) * dim;
        }
        tier = 2;
        return fv.base + nid * dim;
    }
    int32_t g = -1;
    if (fv.slot_of) g = (int32_t)ld_nc_u32(fv.slot_of + nid);
    if (g < 0) { tier = 2; return fv.base + nid * dim; }              // miss -> host / base matrix (Kernels.cu:692-696)
    const int part = (int)(g / fv.cap);                               // hit -> shard[g / cap][g % cap] (Kernels.cu:697-699)
    tier = part == fv.my_part ? 0 : 1;
    return fv.shard_tab[part] + (g - part * fv.cap) * dim;
}

// VEC: 16-byte vectors per lane per row (1 covers D <= 128, 2 covers D <= 256, ...)
template <int VEC, int UNR = GATHER_UNROLL>
__global__ void __launch_bounds__(GATHER_THREADS, UNR == 2 ? 6 : 1) k_gather_v4(const __grid_constant__ FeatView fv, const int32_t* __restrict__ ids,
                                                              const int32_t* __restrict__ nc, int seg_slot, int n_segs,
                                                              float* __restrict__ out, int dim, long long n_nodes,
                                                              long long max_rows, BatchState* __restrict__ st)
{
    const int off = nc[seg_slot];                               // Kernels.cu:672-681; adjacent segments may be fused
    int cnt = 0;
    for (int q = 0; q < n_segs; q++) cnt += nc[seg_slot + 1 + 2 * q];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * GATHER_THREADS) >> 5;
    const int nvec = dim >> 2;
    unsigned long long n_local = 0, n_peer = 0, n_host = 0;
    const unsigned long long stream_pol = policy_evict_first();   // rows are touched once: do not displace the dedup maps
    const unsigned long long keep = policy_evict_last();
    // rows per warp-chunk: 32 for big segments, fewer for small ones so every warp gets work
    int chunk = (cnt + n_warps - 1) / n_warps;
    chunk = chunk >= 32 ? 32 : (chunk <= UNR ? UNR : ((chunk + UNR - 1) / UNR) * UNR);

    for (int r0 = warp * chunk; r0 < cnt; r0 += n_warps * chunk) {
        // lane l: tier + source pointer of row r0+l
        const int r = r0 + lane;
        const uint4* src = nullptr;
        if (lane < chunk && r < cnt && (long long)(off + r) < max_rows) {
            const int32_t id = (int32_t)ld_nc_u32(ids + off + r);
            if (id >= 0) {                                        // Kernels.cu:694
                const long long nid = id < n_nodes ? id : id % n_nodes;
                int tier;
                src = reinterpret_cast<const uint4*>(resolve_row(fv, nid, dim, tier, keep));
                if (tier == 0) n_local++; else if (tier == 1) n_peer++; else n_host++;
            }
        } else if (lane < chunk && r < cnt) {
            st->status = LGN_E_CAPACITY;                          // reference: silent overflow (Server.cu:275)
        }
        const int rows = min(chunk, cnt - r0);
        uint4* dst0 = reinterpret_cast<uint4*>(out + (long long)(off + r0) * dim);
        for (int rr = 0; rr < rows; rr += UNR) {
            uint4 v[UNR][VEC];
            const uint4* sp[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                sp[u] = reinterpret_cast<const uint4*>(__shfl_sync(0xffffffffu, (unsigned long long)src, (rr + u) & 31));
                if (rr + u >= rows) sp[u] = nullptr;
            }
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int k = 0; k < VEC; k++)
                    if (sp[u] && lane + 32 * k < nvec) v[u][k] = ld_stream_v4(sp[u] + lane + 32 * k, stream_pol);
#pragma unroll
            for (int u = 0; u < UNR; u++)
#pragma unroll
                for (int k = 0; k < VEC; k++)
                    if (sp[u] && lane + 32 * k < nvec) st_stream_v4(dst0 + (long long)(rr + u) * nvec + lane + 32 * k, v[u][k], stream_pol);
        }
    }
    // tier statistics for the hit-mix roofline: one atomic per warp per tier
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_local += __shfl_xor_sync(0xffffffffu, n_local, o);
        n_peer += __shfl_xor_sync(0xffffffffu, n_peer, o);
        n_host += __shfl_xor_sync(0xffffffffu, n_host, o);
    }
    if (lane == 0) {
        if (n_local) atomicAdd(&st->tier_rows[0], n_local);
        if (n_peer) atomicAdd(&st->tier_rows[1], n_peer);
        if (n_host) atomicAdd(&st->tier_rows[2], n_host);
    }
}

// ---- bulk-copy (TMA) variant ------------------------------------------------------------
// One thread per row: the thread resolves its row's tier, then the copy engine moves the row
// global -> shared (cp.async.bulk + mbarrier complete_tx) and shared -> global
// (cp.async.bulk ... bulk_group).  Row bytes never pass through registers, a CTA keeps
// blockDim rows in flight with ~30 registers per thread, so the kernel leaves most of each
// SM to the sampling kernels of the next batch that run concurrently.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(GATHER_THREADS) k_gather_bulk(const __grid_constant__ FeatView fv,
                                                                const int32_t* __restrict__ ids,
                                                                const int32_t* __restrict__ nc, int seg_slot, int n_segs,
                                                                float* __restrict__ out, int dim, long long n_nodes,
                                                                long long max_rows, BatchState* __restrict__ st)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int off = nc[seg_slot];
    int cnt = 0;
    for (int q = 0; q < n_segs; q++) cnt += nc[seg_slot + 1 + 2 * q];
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t row_bytes = (uint32_t)dim * 4u;                       // multiple of 16 (checked by the launcher)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)blockDim.x * row_bytes);
    const uint32_t my_buf = smem_u32(smem + (size_t)t * row_bytes);
    const uint32_t my_bar = smem_u32(bars + t);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(my_bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    unsigned long long n_local = 0, n_peer = 0, n_host = 0;
    const unsigned long long stream_pol = policy_evict_first();
    const unsigned long long keep = policy_evict_last();
    uint32_t phase = 0;
    const int stride = gridDim.x * blockDim.x;
    auto resolve = [&](int r) -> const float* {
        if (r >= cnt) return nullptr;
        if ((long long)(off + r) >= max_rows) { st->status = LGN_E_CAPACITY; return nullptr; }
        const int32_t id = (int32_t)ld_nc_u32(ids + off + r);
        if (id < 0) return nullptr;
        const long long nid = id < n_nodes ? id : id % n_nodes;
        int tier;
        const float* p = resolve_row(fv, nid, dim, tier, keep);
        if (tier == 0) n_local++; else if (tier == 1) n_peer++; else n_host++;
        return p;
    };
    int r = blockIdx.x * blockDim.x + t;
    const float* src = resolve(r);
    while (r < cnt) {
        const int rn = r + stride;
        // the bulk store that last read this thread's staging row must have drained
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (src) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(my_bar), "r"(row_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(my_buf), "l"(src), "r"(row_bytes), "r"(my_bar), "l"(stream_pol) : "memory");
        }
        const float* src_next = resolve(rn);        // next row's lookups fly while this row's copy lands
        if (src) {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(my_bar), "r"(phase) : "memory");
            }
            phase ^= 1u;
            LGN_ASSERT((long long)(off + r) < max_rows);
            float* dst = out + (long long)(off + r) * dim;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         ::"l"(dst), "r"(my_buf), "r"(row_bytes), "l"(stream_pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        r = rn;
        src = src_next;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_local += __shfl_xor_sync(0xffffffffu, n_local, o);
        n_peer += __shfl_xor_sync(0xffffffffu, n_peer, o);
        n_host += __shfl_xor_sync(0xffffffffu, n_host, o);
    }
    if (lane == 0) {
        if (n_local) atomicAdd(&st->tier_rows[0], n_local);
        if (n_peer) atomicAdd(&st->tier_rows[1], n_peer);
        if (n_host) atomicAdd(&st->tier_rows[2], n_host);
    }
}

// scalar fallback: D not a multiple of 4 floats or a tier base not 16-byte aligned
__global__ void __launch_bounds__(GATHER_THREADS) k_gather_scalar(const __grid_constant__ FeatView fv, const int32_t* __restrict__ ids,
                                                                  const int32_t* __restrict__ nc, int seg_slot, int n_segs,
                                                                  float* __restrict__ out, int dim, long long n_nodes,
                                                                  long long max_rows, BatchState* __restrict__ st)
{
    const int off = nc[seg_slot];
    int cnt = 0;
    for (int q = 0; q < n_segs; q++) cnt += nc[seg_slot + 1 + 2 * q];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * GATHER_THREADS) >> 5;
    for (int r = warp; r < cnt; r += n_warps) {
        if ((long long)(off + r) >= max_rows) { if (lane == 0) st->status = LGN_E_CAPACITY; continue; }
        const int32_t id = ids[off + r];
        if (id < 0) continue;
        const long long nid = id < n_nodes ? id : id % n_nodes;
        int tier;
        const float* src = resolve_row(fv, nid, dim, tier, policy_evict_last());
        float* dst = out + (long long)(off + r) * dim;
        for (int k = lane; k < dim; k += 32) dst[k] = src[k];
        if (lane == 0) atomicAdd(&st->tier_rows[tier], 1ull);
    }
}

// shard fill: row r of shard j <- src[order[r*kg + j]] (FeatFillUp, GPUCache.cu:200-205)
__global__ void __launch_bounds__(GATHER_THREADS) k_row_copy(const int32_t* __restrict__ order, long long n, long long cap,
                                                             int kg, int j, long long n_repl, const float* __restrict__ src, int dim,
                                                             float* __restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * GATHER_THREADS) >> 5;
    const bool vec = (dim & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0;
    for (long long r = warp; r < cap; r += n_warps) {
        const long long rank = r < n_repl ? r : n_repl + (r - n_repl) * kg + j;   // replicated head, partitioned tail
        if (rank >= n) continue;
        const long long id = order[rank];
        if (vec) {
            const uint4* s = reinterpret_cast<const uint4*>(src + id * dim);
            uint4* d = reinterpret_cast<uint4*>(dst + r * dim);
            for (int k = lane; k < (dim >> 2); k += 32) d[k] = ld_nc_v4(s + k);
        } else {
            for (int k = lane; k < dim; k += 32) dst[r * dim + k] = src[id * dim + k];
        }
    }
}

// diagnostic kernel (lgn_debug_shard_read): the probe's plain LDG loop over the context's own shard mappings
struct DebugTabs { const float* tab[LGN_MAX_PARTS]; int n; };
__device__ __forceinline__ uint32_t debug_mix(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(GATHER_THREADS) k_debug_shard_read(const __grid_constant__ DebugTabs tb, long long rows_per_shard, long long n_rows,
                                                                     int dim, float* __restrict__ out, uint32_t salt)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * GATHER_THREADS) >> 5;
    const int nvec = dim >> 2;
    for (long long r0 = warp * U; r0 < n_rows; r0 += n_warps * U) {
        uint4 v[U][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long r = r0 + u;
            if (r < n_rows) {
                const uint32_t h = debug_mix((uint32_t)r ^ salt);
                const uint4* sp = reinterpret_cast<const uint4*>(tb.tab[(h >> 20) % tb.n] + (long long)(h % (uint32_t)rows_per_shard) * dim);
#pragma unroll
                for (int k = 0; k < 4; k++) if (lane + 32 * k < nvec) v[u][k] = __ldg(sp + lane + 32 * k);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long r = r0 + u;
            if (r < n_rows) {
                uint4* d = reinterpret_cast<uint4*>(out + r * dim);
#pragma unroll
                for (int k = 0; k < 4; k++) if (lane + 32 * k < nvec) d[lane + 32 * k] = v[u][k];
            }
        }
    }
}

void launch_debug_shard_read(lgn_ctx* c, cudaStream_t s, int pipe, long long n_rows, long long rows_per_shard, bool peers_only, uint32_t salt)
{
    DebugTabs tb;
    tb.n = 0;
    for (int i = 0; i < c->feat.n_parts; i++)
        if (!peers_only || i != c->feat.my_part) tb.tab[tb.n++] = c->feat.shard_tab[i];
    if (tb.n == 0) return;
    k_debug_shard_read<<<c->n_sm * 3, GATHER_THREADS, 0, s>>>(tb, rows_per_shard, n_rows, c->cfg.feat_dim, c->pipe[pipe].features, salt);
}

static int gather_auto_mode(const lgn_ctx* c);

static bool gather_vectorisable(const lgn_ctx* c, const Pipe& p)
{
    const int dim = c->cfg.feat_dim;
    bool vec = (dim & 3) == 0 && ((uintptr_t)c->feat.base & 15) == 0 && ((uintptr_t)p.features & 15) == 0;
    for (int i = 0; i < c->feat.n_parts; i++) vec = vec && ((uintptr_t)c->feat.shard_tab[i] & 15) == 0;
    return vec;
}

const char* gather_kernel_name(const lgn_ctx* c)
{
    const Pipe& p = c->pipe[c->cur_pipe];
    const int mode = gather_auto_mode(c);
    if (!gather_vectorisable(c, p) || (c->cfg.feat_dim >> 2) > 128) return "k_gather_scalar";
    return mode == 1 ? "k_gather_bulk (cp.async.bulk feature extraction)" : "k_gather_v4 (128-bit LDG feature extraction)";
}

void launch_gather(lgn_ctx* c, cudaStream_t s, int segment, int n_segs)
{
    Pipe& p = c->pipe[c->cur_pipe];
    const int dim = c->cfg.feat_dim;
    const int seg_slot = 3 + 2 * segment;
    const bool vec = gather_vectorisable(c, p);
    // grid-stride over 32-row chunks.  2 CTAs/SM (16 warps x 4 rows in flight) already saturate HBM when two batches' gathers
    // overlap and leave half of every SM's registers to the sampling chains of the other lanes: papers100M shape, one GPU
    // 0.097 ms/step against 0.104-0.107 with 3-8 CTAs/SM; two GPUs, half of the rows over NVLink 0.177 against 0.191 (3) and
    // 0.183 (4) (profiles/r02aa_mix_sweep.txt, r02af_n2_ldg_ctas.txt)
    const int blocks = c->n_sm * (c->gather_ldg_ctas > 0 ? c->gather_ldg_ctas : 2);
    FeatView fv = c->feat;
    const int nvec = dim >> 2;
    // Timed alone the two variants are equally fast (both reach the copy peak with two launches overlapping); with the
    // sampling chains of the other lanes beside them the LDG variant wins on 512-byte rows (0.097 vs 0.099 ms/step), the
    // bulk-copy variant on 400-byte rows (0.095 vs 0.100): gather_auto_mode
    const int mode = gather_auto_mode(c);
    if (vec && mode == 1) {
        // staging rows + one mbarrier per thread; keep <= ~100 KB per CTA so two CTAs (or the sampler) fit beside it
        int threads = c->gather_threads;
        while (threads > 32 && (size_t)threads * (dim * 4 + 8) > 100 * 1024) threads -= 32;
        const size_t smem = (size_t)threads * (dim * 4 + 8);
        k_gather_bulk<<<c->n_sm * c->gather_ctas_per_sm, threads, smem, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
    } else if (vec && nvec <= 32 && c->gather_unroll == 2)      // experiment knob: 2 rows in flight per warp, 40 registers (slower: r02y)
        k_gather_v4<1, 2><<<blocks, GATHER_THREADS, 0, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
    else if (vec && nvec <= 32)
        k_gather_v4<1><<<blocks, GATHER_THREADS, 0, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
    else if (vec && nvec <= 64)
        k_gather_v4<2><<<blocks, GATHER_THREADS, 0, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
    else if (vec && nvec <= 128)
        k_gather_v4<4><<<blocks, GATHER_THREADS, 0, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
    else
        k_gather_scalar<<<blocks, GATHER_THREADS, 0, s>>>(fv, p.ids, p.nc, seg_slot, n_segs, p.features, dim, c->cfg.n_nodes, c->max_rows, p.state);
}

// gather variant of a context when LGN_GATHER does not force one (measured on one B200, 4 batches in flight):
// rows that fill whole 512-byte warp transactions (D % 128 == 0) and every cache with peer shards take the register-staged
// LDG variant; other row sizes (products' 400-byte rows leave 7 of 32 lanes idle) take the bulk-copy (TMA) variant
static int gather_auto_mode(const lgn_ctx* c)
{
    if (c->gather_mode >= 0) return c->gather_mode;
    return (c->feat.n_parts > 1 || (c->cfg.feat_dim & 127) == 0) ? 0 : 1;
}

static void set_carveout_all(int pct)
{
    cudaFuncSetAttribute(k_gather_bulk, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_gather_v4<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_gather_v4<1, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_gather_v4<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_gather_v4<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_gather_scalar, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    sampler_set_carveout(pct);
}

// per-device function attributes (a process may drive several GPUs): called once per context from lgn_create
void gather_init_device(const lgn_ctx* c)
{
    cudaFuncSetAttribute(k_gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // LGN_CARVEOUT=<percent>: ONE shared-memory / L1 split for every kernel of the pipeline (experiment knob).  The bulk-copy
    // gather needs ~100 KB of shared memory per CTA, the sampling kernels a few KB; with the driver's per-kernel choice the
    // papers100M step takes 0.1085 ms with the bulk-copy gather, 0.0995 with a common 50 % split (profiles/r02aa_mix_sweep.txt),
    // but products' 400-byte rows lose 3 % with it -- so it stays a knob; the default gather for 512-byte rows (LDG) uses no
    // shared memory at all.
    (void)c;
    if (const char* cv = getenv("LGN_CARVEOUT")) set_carveout_all(atoi(cv));
}

void launch_row_copy(const int32_t* order, long long n, long long cap, int kg, int j, long long n_repl, const float* src, int dim,
                     float* dst, int n_sm, cudaStream_t s)
{
    k_row_copy<<<n_sm * 8, GATHER_THREADS, 0, s>>>(order, n, cap, kg, j, n_repl, src, dim, dst);
}

}  // namespace lgn
