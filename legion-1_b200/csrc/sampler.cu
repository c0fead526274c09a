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
//   k_mark     decides the winners (slot_map[dst] == CAND+slot) == new unique nodes and
//              leaves one packed (valid, new) count per 2048-slot tile.
//   k_assign   sums its predecessors' tile counts (parallel, no look-back chain), compacts
//              the valid edges, appends the winners to sampled_ids, writes both local COO
//              indices; the last tile advances the counters (the reference's <<<1,1>>>
//              update_counter).
//              Winners leave their local index in their own slot_dst entry; an in-hop
//              duplicate whose winner is not numbered yet stores -(winner_slot+2) and is
//              patched from that entry later (lazily by the next hop, finally by k_batch_end),
//              so no separate fix-up launch and no second slot_map probe is needed.
#include "context.h"

namespace lgn {

constexpr int SAMPLE_THREADS = 128;
constexpr int SAMPLE_ITEMS = 32;     // frontier items per CTA: small tiles => many CTAs even for hop 1 (B items)
constexpr int RESOLVE_THREADS = 256;
constexpr int RESOLVE_VEC = 8;       // slots per thread, two int4 loads
constexpr int RESOLVE_TILE = RESOLVE_THREADS * RESOLVE_VEC;


// ------------------------------------------------------------------ batch begin
__global__ void __launch_bounds__(256) k_batch_begin(const int32_t* __restrict__ src_ids,
                                                     const int32_t* __restrict__ src_labels, int32_t count,
                                                     int32_t* __restrict__ ids, int32_t* __restrict__ labels,
                                                     const Dedup dd, int32_t* __restrict__ id_h, long long n_nodes,
                                                     int32_t* __restrict__ nc, int32_t* __restrict__ ec,
                                                     BatchState* __restrict__ st, uint32_t step, uint32_t epoch)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gsz = gridDim.x * blockDim.x;
    const unsigned long long keep = policy_evict_last();
    for (int i = gtid; i < count; i += gsz) {
        const int32_t id = src_ids[i];
        ids[i] = id;
        labels[i] = src_labels ? src_labels[i] : -1;
        // Kernels.cu:88-92; duplicate seeds: lowest index wins (the reference races)
        int32_t h = -1;
        if (id >= 0 && id < n_nodes) {
            h = dedup_claim(dd, id, i, keep);
            if (h < 0) st->status = LGN_E_CAPACITY;
        }
        if (dd.bits) id_h[i] = h;
    }
    if (gtid < 16) {   // update_counter(op 0), Kernels.cu:118-127
        nc[gtid] = (gtid == 0 || gtid == 2 || gtid == 4) ? count : 0;
        ec[gtid] = 0;
    }
    if (gtid == 0) {
        HopState h0; h0.n_items = count; h0.item_base = 0; h0.node_base = count; h0.edge_base = 0;
        st->hop[0] = h0;
        st->step = step;
        st->epoch = epoch;
    }
}

// ------------------------------------------------------------------ sample
template <int RNG, bool PRESC>
__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample(const __grid_constant__ TopoView tv, const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ agg_src_ids,
                                                           int32_t* __restrict__ slot_dst, int32_t* __restrict__ slot_h,
                                                           const Dedup dd,
                                                           BatchState* __restrict__ st, int hop, int f,
                                                           unsigned long long seed, uint32_t* __restrict__ topo_hot,
                                                           long long n_nodes)
{
    __shared__ long long s_start[SAMPLE_ITEMS];
    __shared__ const int32_t* s_base[SAMPLE_ITEMS];
    __shared__ int32_t s_deg[SAMPLE_ITEMS];
    __shared__ int32_t s_src[SAMPLE_ITEMS];
    __shared__ uint32_t s_pow[SAMPLE_ITEMS];     // minstd: 48271^(item*f+1)
    __shared__ uint32_t s_ak[256];               // minstd: 48271^k, k < f <= 256
    __shared__ int32_t s_cnt[SAMPLE_ITEMS];      // presampling: edges sampled out of the item

    const HopState hs = st->hop[hop];
    const int F = hs.n_items;
    const int t = threadIdx.x;
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;   // Kernels.cu:368-374
    const uint32_t step = st->step, epoch = st->epoch;
    const unsigned long long keep = policy_evict_last();

    if (RNG == LGN_RNG_MINSTD)
        for (int k = t; k < f; k += SAMPLE_THREADS) s_ak[k] = minstd_pow((unsigned long long)k);
    // the grid is sized for the SMs, not for the worst-case frontier: CTAs stride over the item tiles
    // that actually exist (F is only known on the device)
    for (int item0 = blockIdx.x * SAMPLE_ITEMS; item0 < F; item0 += gridDim.x * SAMPLE_ITEMS) {
    if (t < SAMPLE_ITEMS) {   // phase 1: one thread per frontier item reads its adjacency descriptor once
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
        if (RNG == LGN_RNG_MINSTD) s_pow[t] = deg > 0 ? minstd_pow((unsigned long long)item * f + 1ull) : 0u;
    }
    __syncthreads();

    const int n_items = min(SAMPLE_ITEMS, F - item0);
    const int total = n_items * f;
    const long long slot0 = (long long)item0 * f;
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
                        pick = deg <= f ? k : philox_pick((unsigned long long)(slot0 + s), epoch, (uint32_t)hop, step, seed, deg);
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
                int32_t h = -1;
                if (d >= 0 && d < n_nodes) {                                   // Kernels.cu:411
                    h = dedup_claim(dd, d, CAND + (int32_t)(slot0 + s), keep);
                    if (h < 0) { st->status = LGN_E_CAPACITY; d = -1; }
                    else if (PRESC) atomicAdd(&s_cnt[il[u]], 1);
                } else {
                    d = -1;
                }
                slot_dst[slot0 + s] = d;
                if (dd.bits) slot_h[slot0 + s] = h;
            }
        }
    }
    __syncthreads();
    if (PRESC) {   // one global atomic per frontier item instead of one per edge (Kernels.cu:525)
        if (t < n_items && s_cnt[t] > 0) atomicAdd(&topo_hot[s_src[t]], (uint32_t)s_cnt[t]);
        __syncthreads();
    }
    }   // item tiles
}

// ------------------------------------------------------------------ mark + assign
// packed per-tile counts: valid edges in the low word, new unique nodes in the high word
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long x, unsigned long long* s_red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();                    // s_red may still be read from the previous use
    if (lane == 0) s_red[warp] = x;
    __syncthreads();
    unsigned long long tot = 0;
#pragma unroll
    for (int w = 0; w < RESOLVE_THREADS / 32; w++) tot += s_red[w];
    return tot;
}

__device__ __forceinline__ void load_slots(const int32_t* __restrict__ p, long long base, long long total, int32_t (&d)[RESOLVE_VEC], int32_t fill)
{
    constexpr int V = RESOLVE_VEC;
    if (base + V - 1 < total) {
#pragma unroll
        for (int q = 0; q < V / 4; q++) {
            const int4 v = *reinterpret_cast<const int4*>(p + base + 4 * q);
            d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; j++) d[j] = base + j < total ? p[base + j] : fill;
    }
}

// pass 1 over the hop's slots: who won (slot_map[dst] == CAND + slot)?  Leaves the probed
// slot_map value next to the draw so pass 2 needs no second random access, and one packed
// (valid, new) count per tile -- the cross-tile prefix is then a plain parallel sum in pass 2
// instead of a serial look-back chain.
__global__ void __launch_bounds__(RESOLVE_THREADS) k_mark(const int32_t* __restrict__ slot_dst, const int32_t* __restrict__ slot_h,
                                                          int32_t* __restrict__ slot_val, const Dedup dd,
                                                          const BatchState* __restrict__ st, int hop, int f,
                                                          unsigned long long* __restrict__ tile_cnt)
{
    __shared__ unsigned long long s_red[RESOLVE_THREADS / 32];
    constexpr int V = RESOLVE_VEC;
    const int t = threadIdx.x;
    const HopState hs = st->hop[hop];
    const long long total = (long long)hs.n_items * f;
    const long long n_tiles = total > 0 ? (total + RESOLVE_TILE - 1) / RESOLVE_TILE : 1;
    const unsigned long long keep = policy_evict_last();
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * RESOLVE_TILE + t * V;
        int32_t d[V], v[V], hd[V];
        load_slots(slot_dst, base, total, d, -1);
        if (dd.bits) load_slots(slot_h, base, total, hd, -1);
#pragma unroll
        for (int j = 0; j < V; j++) v[j] = d[j] >= 0 ? dedup_value(dd, dd.bits ? hd[j] : d[j], keep) : EMPTY;
        unsigned long long cnt = 0;
#pragma unroll
        for (int j = 0; j < V; j++) cnt += (d[j] >= 0 ? 1ull : 0ull) + ((d[j] >= 0 && v[j] == CAND + (int32_t)(base + j)) ? (1ull << 32) : 0ull);
        if (base + V - 1 < total) {
#pragma unroll
            for (int q = 0; q < V / 4; q++)
                *reinterpret_cast<int4*>(slot_val + base + 4 * q) = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < V; j++) if (base + j < total) slot_val[base + j] = v[j];
        }
        const unsigned long long tot = block_sum_u64(cnt, s_red);
        if (t == 0) tile_cnt[tile] = tot;
    }
}

// pass 2: compact the valid edges, append the winners to sampled_ids, write both local COO
// indices; the last tile advances the counters (the reference's <<<1,1>>> update_counter).
__global__ void __launch_bounds__(RESOLVE_THREADS) k_assign(
    int32_t* __restrict__ slot_dst, const int32_t* __restrict__ slot_val, const int32_t* __restrict__ slot_h,
    const int32_t* __restrict__ slot_dst_prev,
    int32_t* __restrict__ ids, int32_t* __restrict__ agg_src_ids, int32_t* __restrict__ agg_dst_ids,
    int32_t* __restrict__ agg_src_off, int32_t* __restrict__ agg_dst_off, const Dedup dd, int32_t* __restrict__ id_h,
    int32_t* __restrict__ nc, int32_t* __restrict__ ec, BatchState* __restrict__ st, int hop, int f,
    const unsigned long long* __restrict__ tile_cnt, long long capacity)
{
    __shared__ unsigned long long s_red[RESOLVE_THREADS / 32];
    __shared__ uint32_t s_warp[RESOLVE_THREADS / 32];
    __shared__ int32_t s_edge[4][RESOLVE_TILE];   // 32 KB: src id, dst id, src local index, dst local index
    __shared__ int32_t s_new[RESOLVE_TILE];       //  8 KB: new unique ids of the tile
    constexpr int V = RESOLVE_VEC;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const HopState hs = st->hop[hop];
    const long long total = (long long)hs.n_items * f;
    const long long n_tiles = total > 0 ? (total + RESOLVE_TILE - 1) / RESOLVE_TILE : 1;
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;
    const unsigned long long keep = policy_evict_last();
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // exclusive prefix of the tile counts: every CTA sums its predecessors in parallel
        unsigned long long part = 0;
        for (long long i = t; i < tile; i += RESOLVE_THREADS) part += tile_cnt[i];
        const unsigned long long prefix = block_sum_u64(part, s_red);

        const long long base = tile * RESOLVE_TILE + t * V;
        int32_t d[V], v[V], hd[V];
        load_slots(slot_dst, base, total, d, -1);
        load_slots(slot_val, base, total, v, EMPTY);
        if (dd.bits) load_slots(slot_h, base, total, hd, -1);
        bool isnew[V];
        uint32_t cnt = 0;   // valid count | new count << 16
#pragma unroll
        for (int j = 0; j < V; j++) {
            isnew[j] = d[j] >= 0 && v[j] == CAND + (int32_t)(base + j);
            cnt += (d[j] >= 0 ? 1u : 0u) + (isnew[j] ? 0x10000u : 0u);
        }
        uint32_t inc = cnt;   // block-exclusive scan of the packed counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        __syncthreads();
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
        const long long e_base = hs.edge_base + (long long)(prefix & 0xffffffffull);
        const long long n_base = hs.node_base + (long long)(prefix >> 32);
        int e_loc = (int)(excl & 0xffffu), n_loc = (int)(excl >> 16);
        // stage the tile's compacted edges / new ids in shared memory, then write them out coalesced
#pragma unroll
        for (int j = 0; j < V; j++) {
            if (d[j] < 0) continue;
            const int item = (int)((base + j) / f);
            const int32_t src = frontier[item];
            // local index of the frontier node: a seed's index sits in slot_map, a later hop's
            // item is the previous hop's edge, already relabelled (construct_graph, Kernels.cu:458-461)
            int32_t dst_off = hop == 0 ? dedup_value(dd, dd.bits ? id_h[item] : src, keep) : agg_src_off[hs.item_base + item];
            if (dst_off < 0) dst_off = slot_dst_prev[-2 - dst_off];   // duplicate of the previous hop: its winner's entry
            int32_t src_off;
            if (isnew[j]) {                       // Kernels.cu:418-438
                const long long pos = n_base + n_loc;
                s_new[n_loc++] = d[j];
                if (pos < capacity) {
                    dedup_publish(dd, dd.bits ? hd[j] : d[j], d[j], (int32_t)pos, keep);
                    if (dd.bits) id_h[pos] = hd[j];
                }
                slot_dst[base + j] = (int32_t)pos;   // winners publish their local index in their own slot
                src_off = (int32_t)pos;
            } else {
                src_off = v[j] < CAND ? v[j] : -2 - (v[j] - CAND);   // winner of this hop not numbered yet: remember its slot
            }
            s_edge[0][e_loc] = d[j];             // Kernels.cu:423-424, 441-445
            s_edge[1][e_loc] = src;
            s_edge[2][e_loc] = src_off;
            s_edge[3][e_loc] = dst_off;
            e_loc++;
        }
        __syncthreads();
        {
            const int n_valid = (int)(block_tot & 0xffffu), n_new_blk = (int)(block_tot >> 16);
            for (int i = t; i < n_valid; i += RESOLVE_THREADS) {
                const long long e = e_base + i;
                if (e < capacity) {
                    agg_src_ids[e] = s_edge[0][i];
                    agg_dst_ids[e] = s_edge[1][i];
                    agg_src_off[e] = s_edge[2][i];
                    agg_dst_off[e] = s_edge[3][i];
                }
            }
            for (int i = t; i < n_new_blk; i += RESOLVE_THREADS)
                if (n_base + i < capacity) ids[n_base + i] = s_new[i];
        }
        if (tile == n_tiles - 1 && t == 0) {      // update_counter(op 2/4), Kernels.cu:128-149, any hop
            const int32_t n_e = (int32_t)(prefix & 0xffffffffull) + (int32_t)(block_tot & 0xffffu);
            const int32_t n_new = (int32_t)(prefix >> 32) + (int32_t)(block_tot >> 16);
            HopState nx;
            nx.n_items = n_e; nx.item_base = hs.edge_base; nx.node_base = hs.node_base + n_new; nx.edge_base = hs.edge_base + n_e;
            st->hop[hop + 1] = nx;
            st->tot_items += (unsigned long long)hs.n_items;
            st->tot_edges += (unsigned long long)n_e;
            if ((long long)nx.node_base > capacity || (long long)nx.edge_base > capacity) st->status = LGN_E_CAPACITY;
            nc[0] = nx.node_base; nc[1] = 0; nc[2] = n_e;
            nc[5 + 2 * hop] = hs.node_base; nc[6 + 2 * hop] = n_new;
            if (7 + 2 * hop < 16) nc[7 + 2 * hop] = nx.node_base;
            ec[0] = nx.edge_base; ec[1] = 0; ec[2] = hs.edge_base; ec[3 + hop] = nx.edge_base;
        }
    }
}

// ------------------------------------------------------------------ batch end
struct SlotRegions { long long off[LGN_MAX_HOPS + 1]; };

template <bool PRESC>
__global__ void __launch_bounds__(256) k_batch_end(const int32_t* __restrict__ ids, const Dedup dd, const int32_t* __restrict__ id_h,
                                                   BatchState* __restrict__ st, int n_hops,
                                                   uint32_t* __restrict__ node_hot, long long n_nodes,
                                                   int32_t* __restrict__ agg_src_off, const int32_t* __restrict__ slot_dst,
                                                   const __grid_constant__ SlotRegions reg)
{
    // patch the in-hop duplicates left by k_resolve (construct_graph's second lookup, Kernels.cu:458)
    const int n_edges = st->hop[n_hops].edge_base;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
        const int32_t so = agg_src_off[e];
        if (so < 0) {
            int h = 0;
            while (h + 1 < n_hops && e >= st->hop[h + 1].edge_base) h++;
            agg_src_off[e] = slot_dst[reg.off[h] + (-2 - so)];
        }
    }
    const int total = st->hop[n_hops].node_base;   // nc[9] for two hops
    const unsigned long long keep = policy_evict_last();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int32_t id = ids[i];
        if (id >= 0 && id < n_nodes) {
            if (PRESC) atomicAdd(&node_hot[id], 1u);   // HotnessMeasure: ids of a batch are unique -> no contention
            const int32_t h = dd.bits ? id_h[i] : id;
            if (h >= 0) dedup_release(dd, h, keep);     // ClearPosMap + the reference's per-batch N/8-byte memset
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
    step += c->rng_step_offset;
    int blocks = cdiv(count, 256);
    if (blocks < 1) blocks = 1;
    k_batch_begin<<<blocks, 256, 0, s>>>(ids + src_off, labels ? labels + src_off : nullptr, count, p.ids, p.labels,
                                         p.dedup, p.id_h, c->cfg.n_nodes, p.nc, p.ec, p.state, step, c->rng_epoch);
}

void launch_sample_hop(lgn_ctx* c, cudaStream_t s, int hop, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    long long fmax = c->cfg.batch_size;
    for (int h = 0; h < hop; h++) fmax *= c->cfg.fanout[h];
    const int f = c->cfg.fanout[hop];
    int sblocks = cdiv(fmax, SAMPLE_ITEMS);
    if (sblocks > c->n_sm * c->sample_ctas_per_sm) sblocks = c->n_sm * c->sample_ctas_per_sm;
    int32_t* slot_dst = p.slot_dst + c->slot_off[hop];
    const int32_t* slot_prev = hop > 0 ? p.slot_dst + c->slot_off[hop - 1] : p.slot_dst;
#define LGN_SAMPLE(R, P)                                                                                     \
    k_sample<R, P><<<sblocks, SAMPLE_THREADS, 0, s>>>(c->topo, p.ids, p.agg_src_ids, slot_dst, p.slot_h, p.dedup, p.state, \
                                                      hop, f, c->cfg.rng_seed, c->topo_hotness, c->cfg.n_nodes)
    if (c->cfg.rng_mode == LGN_RNG_MINSTD) { if (presc) LGN_SAMPLE(LGN_RNG_MINSTD, true); else LGN_SAMPLE(LGN_RNG_MINSTD, false); }
    else { if (presc) LGN_SAMPLE(LGN_RNG_PHILOX, true); else LGN_SAMPLE(LGN_RNG_PHILOX, false); }
#undef LGN_SAMPLE
    int rblocks = cdiv(fmax * f, RESOLVE_TILE) + 1;
    if (rblocks > c->n_sm * c->resolve_ctas_per_sm) rblocks = c->n_sm * c->resolve_ctas_per_sm;
    k_mark<<<rblocks, RESOLVE_THREADS, 0, s>>>(slot_dst, p.slot_h, p.slot_val, p.dedup, p.state, hop, f, p.tile_cnt);
    k_assign<<<rblocks, RESOLVE_THREADS, 0, s>>>(slot_dst, p.slot_val, p.slot_h, slot_prev, p.ids, p.agg_src_ids, p.agg_dst_ids,
                                                 p.agg_src_off, p.agg_dst_off, p.dedup, p.id_h, p.nc, p.ec, p.state, hop, f,
                                                 p.tile_cnt, c->capacity);
}

void launch_batch_end(lgn_ctx* c, cudaStream_t s, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    int blocks = cdiv(c->capacity, 256);
    if (blocks > c->n_sm * c->end_ctas_per_sm) blocks = c->n_sm * c->end_ctas_per_sm;
    SlotRegions reg;
    for (int h = 0; h <= LGN_MAX_HOPS; h++) reg.off[h] = c->slot_off[h];
    if (presc) k_batch_end<true><<<blocks, 256, 0, s>>>(p.ids, p.dedup, p.id_h, p.state, c->cfg.n_hops, c->node_hotness, c->cfg.n_nodes, p.agg_src_off, p.slot_dst, reg);
    else k_batch_end<false><<<blocks, 256, 0, s>>>(p.ids, p.dedup, p.id_h, p.state, c->cfg.n_hops, nullptr, c->cfg.n_nodes, p.agg_src_off, p.slot_dst, reg);
}

}  // namespace lgn
