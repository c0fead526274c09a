// sampler.cu -- batch generation, k-hop neighbour sampling, dedup/relabel and
// presampling hotness for sm_100a.
//
// Replaces (reference paths): batch_generator + update_counter (Kernels.cu:68-150),
// kernel_random_sampler_2 / kernel_pre_sampler_optimized (Kernels.cu:342-564),
// construct_graph (Kernels.cu:450-463), ClearPosMap (Kernels.cu:750-756),
// HotnessMeasure (GPUCache.cu:227-235) and the two cuckoo lookups of FindTopo
// (GPUCache.cu:434-443).
//
// Design: three launches per hop, no host synchronisation,
// deterministic output order ("first occurrence in slot order"):
//   k_sample   one CTA per 256 frontier items: the item's thread reads its indptr
//              pair once (the reference re-reads it in each of the `fanout` threads),
//              then the CTA's lanes sweep the items' fanout slots, draw (Philox4x32-10
//              or the minstd closed form), read the neighbour id and atomicMin the
//              slot index into slot_map[dst].  Draw results land uncompacted in slot_dst.
//   k_resolve  single-pass decoupled look-back scan over the slots: compacts the valid
//              edges, decides the winners (slot_map[dst] == CAND+slot) == new unique nodes,
//              appends them to sampled_ids, writes both local COO indices; the last tile
//              advances the counters (the reference's <<<1,1>>> update_counter).
//   k_fix      patches the local index of in-hop duplicates whose winner had not been
//              numbered yet when they were visited.
#include "context.h"

namespace lgn {

constexpr int SAMPLE_THREADS = 256;
constexpr int RESOLVE_THREADS = 256;
constexpr int RESOLVE_TILE = 1024;   // 4 slots per thread, one int4 load

constexpr unsigned long long FLAG_A = 1ull << 62, FLAG_P = 2ull << 62, FLAG_MASK = 3ull << 62;
constexpr unsigned long long FIELD = 0x7fffffffull;

// ------------------------------------------------------------------ batch begin
__global__ void __launch_bounds__(256) k_batch_begin(const int32_t* __restrict__ src_ids,
                                                     const int32_t* __restrict__ src_labels, int32_t count,
                                                     int32_t* __restrict__ ids, int32_t* __restrict__ labels,
                                                     int32_t* __restrict__ slot_map, long long n_nodes,
                                                     int32_t* __restrict__ nc, int32_t* __restrict__ ec,
                                                     BatchState* __restrict__ st, uint32_t step,
                                                     unsigned long long* __restrict__ scan_status, int n_status,
                                                     int32_t* __restrict__ scan_ticket)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsz = gridDim.x * blockDim.x;
    for (int i = gtid; i < count; i += gsz) {
        const int32_t id = src_ids[i];
        ids[i] = id;
        labels[i] = src_labels ? src_labels[i] : -1;
        // Kernels.cu:88-92; duplicate seeds: lowest index wins (the reference races)
        if (id >= 0 && id < n_nodes) atomicMin(&slot_map[id], i);
    }
    for (int i = gtid; i < n_status; i += gsz) scan_status[i] = 0ull;
    if (gtid < LGN_MAX_HOPS) scan_ticket[gtid] = 0;
    if (gtid < 16) {   // update_counter(op 0), Kernels.cu:118-127
        nc[gtid] = (gtid == 0 || gtid == 2 || gtid == 4) ? count : 0;
        ec[gtid] = 0;
    }
    if (gtid == 0) {
        HopState h0; h0.n_items = count; h0.item_base = 0; h0.node_base = count; h0.edge_base = 0;
        st->hop[0] = h0;
        st->step = step;
    }
}

// ------------------------------------------------------------------ sample
template <int RNG, bool PRESC>
__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample(const __grid_constant__ TopoView tv, const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ agg_src_ids,
                                                           int32_t* __restrict__ slot_dst,
                                                           int32_t* __restrict__ slot_map,
                                                           const BatchState* __restrict__ st, int hop, int f,
                                                           unsigned long long seed, uint32_t* __restrict__ topo_hot,
                                                           long long n_nodes)
{
    __shared__ long long s_start[SAMPLE_THREADS];
    __shared__ const int32_t* s_base[SAMPLE_THREADS];
    __shared__ int32_t s_deg[SAMPLE_THREADS];
    __shared__ int32_t s_src[SAMPLE_THREADS];
    __shared__ uint32_t s_pow[SAMPLE_THREADS];   // minstd: 48271^(item*f+1)
    __shared__ uint32_t s_ak[SAMPLE_THREADS];    // minstd: 48271^k, k < f
    __shared__ int32_t s_cnt[SAMPLE_THREADS];    // presampling: edges sampled out of the item

    const HopState hs = st->hop[hop];
    const int F = hs.n_items;
    const int item0 = blockIdx.x * SAMPLE_THREADS;
    if (item0 >= F) return;
    const int t = threadIdx.x;
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;   // Kernels.cu:368-374

    {   // phase 1: one thread per frontier item reads its adjacency descriptor once
        const int item = item0 + t;
        int32_t src = -1, deg = 0;
        long long start = 0;
        const int32_t* base = tv.base_indices;
        if (item < F) {
            src = frontier[item];
            if (src >= 0 && src < n_nodes) {                                 // Kernels.cu:385
                int32_t g = -1;
                if (!PRESC && tv.slot_of) g = (int32_t)ld_nc_u32(tv.slot_of + src);
                const int64_t* ip;
                if (g < 0) {                                                   // miss: base CSR (Kernels.cu:391-393)
                    ip = tv.base_indptr + src;
                } else {                                                       // hit: local or peer shard (:394-396)
                    const int part = (int)(g / tv.cap);
                    ip = tv.indptr_tab[part] + (g - part * tv.cap);
                    base = tv.indices_tab[part];
                }
                start = ip[0];
                deg = (int32_t)(ip[1] - start);
            } else {
                src = -1;
            }
        }
        s_start[t] = start; s_base[t] = base; s_deg[t] = deg; s_src[t] = src;
        if (PRESC) s_cnt[t] = 0;
        if (RNG == LGN_RNG_MINSTD) {
            s_pow[t] = deg > 0 ? minstd_pow((unsigned long long)item * f + 1ull) : 0u;
            if (t < f) s_ak[t] = minstd_pow((unsigned long long)t);
        }
    }
    __syncthreads();

    const int n_items = min(SAMPLE_THREADS, F - item0);
    const int total = n_items * f;
    const long long slot0 = (long long)item0 * f;
    const uint32_t step = st->step;
    constexpr int U = 4;
    for (int s0 = t; s0 < total; s0 += SAMPLE_THREADS * U) {
        int32_t dst[U];
        int il[U];
#pragma unroll
        for (int u = 0; u < U; u++) {   // issue the U neighbour reads back to back
            const int s = s0 + u * SAMPLE_THREADS;
            dst[u] = -1;
            il[u] = 0;
            if (s < total) {
                const int i = s / f, k = s - i * f;
                il[u] = i;
                const int deg = s_deg[i];
                if (k < deg) {                                                 // Kernels.cu:399
                    int32_t pick;
                    if (RNG == LGN_RNG_MINSTD) {
                        pick = minstd_to_pick(mulmod_m31(s_pow[i], s_ak[k]), deg);
                    } else {
                        pick = deg <= f ? k : philox_pick((unsigned long long)(slot0 + s), (uint32_t)hop, step, seed, deg);
                    }
                    dst[u] = (int32_t)ld_nc_u32(s_base[i] + s_start[i] + pick);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = s0 + u * SAMPLE_THREADS;
            if (s < total) {
                int32_t d = dst[u];
                if (d >= 0 && d < n_nodes) {                                   // Kernels.cu:411
                    atomicMin(&slot_map[d], CAND + (int32_t)(slot0 + s));
                    if (PRESC) atomicAdd(&s_cnt[il[u]], 1);
                } else {
                    d = -1;
                }
                slot_dst[slot0 + s] = d;
            }
        }
    }
    if (PRESC) {   // one global atomic per frontier item instead of one per edge (Kernels.cu:525)
        __syncthreads();
        if (t < n_items && s_cnt[t] > 0) atomicAdd(&topo_hot[s_src[t]], (uint32_t)s_cnt[t]);
    }
}

// ------------------------------------------------------------------ resolve
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(RESOLVE_THREADS) k_resolve(
    const int32_t* __restrict__ slot_dst, int32_t* __restrict__ ids, int32_t* __restrict__ agg_src_ids,
    int32_t* __restrict__ agg_dst_ids, int32_t* __restrict__ agg_src_off, int32_t* __restrict__ agg_dst_off,
    int32_t* __restrict__ slot_map, int32_t* __restrict__ nc, int32_t* __restrict__ ec, BatchState* __restrict__ st,
    int hop, int f, unsigned long long* __restrict__ status, int32_t* __restrict__ ticket, long long capacity)
{
    __shared__ int s_tile;
    __shared__ uint32_t s_warp[RESOLVE_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) s_tile = atomicAdd(ticket, 1);   // ticket order => every lower tile is already running
    __syncthreads();
    const int tile = s_tile;
    const HopState hs = st->hop[hop];
    const long long total = (long long)hs.n_items * f;
    const long long n_tiles = total > 0 ? (total + RESOLVE_TILE - 1) / RESOLVE_TILE : 1;
    if (tile >= n_tiles) return;

    const long long base = (long long)tile * RESOLVE_TILE + t * 4;
    int32_t d[4] = {-1, -1, -1, -1};
    if (base + 3 < total) {
        const int4 v = *reinterpret_cast<const int4*>(slot_dst + base);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (base + j < total) d[j] = slot_dst[base + j];
    }
    int32_t v[4];
    uint32_t cnt = 0;   // valid count | new count << 16
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = d[j] >= 0 ? slot_map[d[j]] : EMPTY;
    bool isnew[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        isnew[j] = d[j] >= 0 && v[j] == CAND + (int32_t)(base + j);
        cnt += (d[j] >= 0 ? 1u : 0u) + (isnew[j] ? 0x10000u : 0u);
    }
    // block-exclusive scan of the packed counts
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, block_tot = 0;
#pragma unroll
    for (int w = 0; w < RESOLVE_THREADS / 32; w++) {
        const uint32_t x = s_warp[w];
        if (w < warp) warp_off += x;
        block_tot += x;
    }
    const uint32_t excl = warp_off + inc - cnt;

    // decoupled look-back across tiles, warp-parallel
    if (warp == 0) {
        const unsigned long long agg = (unsigned long long)(block_tot & 0xffffu) | ((unsigned long long)(block_tot >> 16) << 31);
        unsigned long long run = 0;
        if (tile > 0) {
            if (lane == 0) st_status(status + tile, FLAG_A | agg);
            long long p = (long long)tile - 1;
            while (true) {
                const long long q = p - lane;
                unsigned long long w = q >= 0 ? ld_status(status + q) : FLAG_P;
                while (__any_sync(0xffffffffu, (w & FLAG_MASK) == 0ull)) w = q >= 0 ? ld_status(status + q) : FLAG_P;
                const unsigned pm = __ballot_sync(0xffffffffu, (w & FLAG_MASK) == FLAG_P);
                const int first = pm ? __ffs(pm) - 1 : 31;
                unsigned long long val = lane <= first ? (w & ~FLAG_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                run += val;
                if (pm) break;
                p -= 32;
            }
        }
        if (lane == 0) {
            st_status(status + tile, FLAG_P | (run + agg));
            s_prefix = run;
        }
    }
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    long long e_cur = hs.edge_base + (long long)(prefix & FIELD) + (excl & 0xffffu);
    long long n_cur = hs.node_base + (long long)((prefix >> 31) & FIELD) + (excl >> 16);
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (d[j] < 0) continue;
        const int item = (int)((base + j) / f);
        const int32_t src = frontier[item];
        // local index of the frontier node: a seed's index sits in slot_map, a later hop's
        // item is the previous hop's edge, already relabelled (construct_graph, Kernels.cu:458-461)
        const int32_t dst_off = hop == 0 ? slot_map[src] : agg_src_off[hs.item_base + item];
        int32_t src_off;
        if (isnew[j]) {                       // Kernels.cu:418-438
            const long long pos = n_cur++;
            if (pos < capacity) { ids[pos] = d[j]; slot_map[d[j]] = (int32_t)pos; }
            src_off = (int32_t)pos;
        } else {
            src_off = v[j] < CAND ? v[j] : -1;   // -1: winner of this hop not numbered yet -> k_fix
        }
        const long long e = e_cur++;
        if (e < capacity) {                  // Kernels.cu:423-424, 441-445
            agg_src_ids[e] = d[j];
            agg_dst_ids[e] = src;
            agg_src_off[e] = src_off;
            agg_dst_off[e] = dst_off;
        }
    }
    if (tile == n_tiles - 1 && t == 0) {      // update_counter(op 2/4), Kernels.cu:128-149, any hop
        const int32_t n_e = (int32_t)(prefix & FIELD) + (int32_t)(block_tot & 0xffffu);
        const int32_t n_new = (int32_t)((prefix >> 31) & FIELD) + (int32_t)(block_tot >> 16);
        HopState nx;
        nx.n_items = n_e; nx.item_base = hs.edge_base; nx.node_base = hs.node_base + n_new; nx.edge_base = hs.edge_base + n_e;
        st->hop[hop + 1] = nx;
        if ((long long)nx.node_base > capacity || (long long)nx.edge_base > capacity) st->status = LGN_E_CAPACITY;
        nc[0] = nx.node_base; nc[1] = 0; nc[2] = n_e;
        nc[5 + 2 * hop] = hs.node_base; nc[6 + 2 * hop] = n_new;
        if (7 + 2 * hop < 16) nc[7 + 2 * hop] = nx.node_base;
        ec[0] = nx.edge_base; ec[1] = 0; ec[2] = hs.edge_base; ec[3 + hop] = nx.edge_base;
    }
}

__global__ void __launch_bounds__(256) k_fix(const int32_t* __restrict__ agg_src_ids, int32_t* __restrict__ agg_src_off,
                                             const int32_t* __restrict__ slot_map, const BatchState* __restrict__ st, int hop)
{
    const int lo = st->hop[hop].edge_base, hi = st->hop[hop + 1].edge_base;
    for (int e = lo + blockIdx.x * blockDim.x + threadIdx.x; e < hi; e += gridDim.x * blockDim.x)
        if (agg_src_off[e] < 0) agg_src_off[e] = slot_map[agg_src_ids[e]];
}

// ------------------------------------------------------------------ batch end
template <bool PRESC>
__global__ void __launch_bounds__(256) k_batch_end(const int32_t* __restrict__ ids, int32_t* __restrict__ slot_map,
                                                   BatchState* __restrict__ st, int n_hops,
                                                   uint32_t* __restrict__ node_hot, long long n_nodes)
{
    const int total = st->hop[n_hops].node_base;   // nc[9] for two hops
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int32_t id = ids[i];
        if (id >= 0 && id < n_nodes) {
            if (PRESC) atomicAdd(&node_hot[id], 1u);   // HotnessMeasure: ids of a batch are unique -> no contention
            slot_map[id] = EMPTY;                       // ClearPosMap + the reference's per-batch N/8-byte memset
        }
    }
    if (PRESC && blockIdx.x == 0 && threadIdx.x == 0 && total > st->max_ids) st->max_ids = total;
}

// ------------------------------------------------------------------ launchers
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

void launch_batch_begin(lgn_ctx* c, cudaStream_t s, const int32_t* ids, const int32_t* labels, int32_t src_off,
                        int32_t count, uint32_t step)
{
    Pipe& p = c->pipe[c->cur_pipe];
    const int n_status = cdiv(c->max_slots, RESOLVE_TILE) + 1;
    int blocks = cdiv(count > n_status ? count : n_status, 256);
    if (blocks < 1) blocks = 1;
    k_batch_begin<<<blocks, 256, 0, s>>>(ids + src_off, labels ? labels + src_off : nullptr, count, p.ids, p.labels,
                                         c->slot_map, c->cfg.n_nodes, p.nc, p.ec, c->state, step, c->scan_status,
                                         n_status * LGN_MAX_HOPS, c->scan_ticket);
}

void launch_sample_hop(lgn_ctx* c, cudaStream_t s, int hop, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    long long fmax = c->cfg.batch_size;
    for (int h = 0; h < hop; h++) fmax *= c->cfg.fanout[h];
    const int f = c->cfg.fanout[hop];
    const int n_status = cdiv(c->max_slots, RESOLVE_TILE) + 1;
    const int sblocks = cdiv(fmax, SAMPLE_THREADS);
    TopoView tv = c->topo;
#define LGN_SAMPLE(R, P)                                                                                          \
    k_sample<R, P><<<sblocks, SAMPLE_THREADS, 0, s>>>(tv, p.ids, c->agg_src_ids, c->slot_dst, c->slot_map, c->state, \
                                                      hop, f, c->cfg.rng_seed, c->topo_hotness, c->cfg.n_nodes)
    if (c->cfg.rng_mode == LGN_RNG_MINSTD) { if (presc) LGN_SAMPLE(LGN_RNG_MINSTD, true); else LGN_SAMPLE(LGN_RNG_MINSTD, false); }
    else { if (presc) LGN_SAMPLE(LGN_RNG_PHILOX, true); else LGN_SAMPLE(LGN_RNG_PHILOX, false); }
#undef LGN_SAMPLE
    const int rblocks = cdiv(fmax * f, RESOLVE_TILE) + 1;
    k_resolve<<<rblocks, RESOLVE_THREADS, 0, s>>>(c->slot_dst, p.ids, c->agg_src_ids, c->agg_dst_ids, p.agg_src_off,
                                                  p.agg_dst_off, c->slot_map, p.nc, p.ec, c->state, hop, f,
                                                  c->scan_status + (size_t)hop * n_status, c->scan_ticket + hop,
                                                  c->capacity);
    int fblocks = cdiv(fmax * f, 256);
    if (fblocks > c->n_sm * 8) fblocks = c->n_sm * 8;
    k_fix<<<fblocks, 256, 0, s>>>(c->agg_src_ids, p.agg_src_off, c->slot_map, c->state, hop);
}

void launch_batch_end(lgn_ctx* c, cudaStream_t s, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    int blocks = cdiv(c->capacity, 256);
    if (blocks > c->n_sm * 8) blocks = c->n_sm * 8;
    if (presc) k_batch_end<true><<<blocks, 256, 0, s>>>(p.ids, c->slot_map, c->state, c->cfg.n_hops, c->node_hotness, c->cfg.n_nodes);
    else k_batch_end<false><<<blocks, 256, 0, s>>>(p.ids, c->slot_map, c->state, c->cfg.n_hops, nullptr, c->cfg.n_nodes);
}

}  // namespace lgn
