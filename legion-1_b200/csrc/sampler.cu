// sampler.cu -- batch generation, k-hop neighbour sampling, dedup/relabel and
// presampling hotness for sm_100a.
//
// Replaces (reference paths): batch_generator + update_counter (Kernels.cu:68-150),
// kernel_random_sampler_2 / kernel_pre_sampler_optimized (Kernels.cu:342-564),
// construct_graph (Kernels.cu:450-463), ClearPosMap (Kernels.cu:750-756),
// HotnessMeasure (GPUCache.cu:227-235) and the two cuckoo lookups of FindTopo
// (GPUCache.cu:434-443).
//
// Design: three launches per hop, no host synchronisation, deterministic output
// order ("first occurrence in slot order"), and every pass after the draw touches only the VALID draws:
//   k_sample   one CTA per tile of 32 frontier items: the item's thread reads its indptr pair once (the
//              reference re-reads it in each of the `fanout` threads), the CTA's lanes sweep the tile's
//              32*f slots: draw (Philox4x32-10 or the minstd closed form), read the neighbour id, red.min
//              "generation | CAND | slot" into the node's dedup entry.  The valid draws of the tile are then
//              compacted in slot order (shared-memory pass) into the tile's own region: later passes read
//              e entries, not F*f slots.
//   k_mark     one warp per tile, 8 entries per lane in flight: probes each draw's entry once.  Winner <=> the entry
//              still holds this draw's slot, i.e. it is the first occurrence of a node not seen in earlier hops.
//              Leaves the probed payload (and the key) next to the draw so the next pass needs no random access,
//              and (edges, new nodes) counts per tile and per 64 tiles.
//   k_assign   one warp per tile: exclusive prefix from the per-64-tile sums + the tile counts of its own group (no
//              device-wide scan, no look-back chain), edge index = prefix + position in the tile, winners numbered
//              by ballot scans, appended to sampled_ids and published in their dedup entry; both local COO indices
//              are written.  An in-hop duplicate whose winner is numbered by another thread stores -(handle+2);
//              the next hop reads such an index through the dedup entry, k_batch_end patches what is left.  The
//              warp of the last tile advances the batch counters (the reference's <<<1,1>>> update_counter).
// Nothing is released at the end of a batch: dedup values carry a generation (common.cuh).
#include "context.h"

namespace lgn {

constexpr int SAMPLE_THREADS = 128;
constexpr int SAMPLE_ITEMS = 32;     // frontier items per tile: small tiles => many CTAs even for hop 1 (B items)
constexpr int SAMPLE_WARPS = SAMPLE_THREADS / 32;
constexpr int SCAN_THREADS = 128;    // k_mark / k_assign CTA size (one tile per WARP)

// ------------------------------------------------------------------ batch begin
__global__ void __launch_bounds__(256) k_batch_begin(const int32_t* __restrict__ src_ids,
                                                     const int32_t* __restrict__ src_labels, int32_t count,
                                                     int32_t* __restrict__ ids, int32_t* __restrict__ labels,
                                                     const Dedup dd, int32_t* __restrict__ seed_h, long long n_nodes,
                                                     int32_t* __restrict__ nc, int32_t* __restrict__ ec,
                                                     BatchState* __restrict__ st, uint32_t step, uint32_t epoch, int32_t gen_base)
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
            h = dedup_claim(dd, id, gen_base | i, keep);
            if (h < 0) st->status = LGN_E_CAPACITY;
        }
        if (dd.bits) seed_h[i] = h;
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
        st->gen_base = gen_base;
    }
}

// ------------------------------------------------------------------ sample
// dynamic shared memory: int32 s_val[SAMPLE_ITEMS * f] -- the tile's draws (dedup handle, -1 = no edge)
template <int RNG, bool PRESC>
__global__ void __launch_bounds__(SAMPLE_THREADS) k_sample(const __grid_constant__ TopoView tv, const int32_t* __restrict__ ids,
                                                           const int32_t* __restrict__ agg_src_ids,
                                                           int32_t* __restrict__ draw_h, uint16_t* __restrict__ draw_s,
                                                           int32_t* __restrict__ tile_n, int32_t* __restrict__ super_e,
                                                           int32_t* __restrict__ super_n, const Dedup dd,
                                                           BatchState* __restrict__ st, int hop, int f,
                                                           unsigned long long seed, uint32_t* __restrict__ topo_hot,
                                                           long long n_nodes)
{
    extern __shared__ int32_t s_val[];
    __shared__ long long s_start[SAMPLE_ITEMS];
    __shared__ const int32_t* s_base[SAMPLE_ITEMS];
    __shared__ int32_t s_deg[SAMPLE_ITEMS];
    __shared__ int32_t s_src[SAMPLE_ITEMS];
    __shared__ uint32_t s_pow[SAMPLE_ITEMS];     // minstd: 48271^(item*f+1)
    __shared__ uint32_t s_ak[256];               // minstd: 48271^k, k < f <= 256
    __shared__ int32_t s_cnt[SAMPLE_ITEMS];      // presampling: edges sampled out of the item
    __shared__ int32_t s_wcnt[SAMPLE_WARPS];

    const HopState hs = st->hop[hop];
    const int F = hs.n_items;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;   // Kernels.cu:368-374
    const uint32_t step = st->step, epoch = st->epoch;
    const int32_t cand_base = st->gen_base | CAND;
    const unsigned long long keep = policy_evict_last();

    if (RNG == LGN_RNG_MINSTD)
        for (int k = t; k < f; k += SAMPLE_THREADS) s_ak[k] = minstd_pow((unsigned long long)k);
    // the grid is sized for the SMs, not for the worst-case frontier: CTAs stride over the item tiles
    // that actually exist (F is only known on the device)
    const int n_tiles = (F + SAMPLE_ITEMS - 1) / SAMPLE_ITEMS;
    if (blockIdx.x == 0)      // per-supertile sums of the next pass start from zero
        for (int i = t; i <= n_tiles / 64; i += SAMPLE_THREADS) { super_e[i] = 0; super_n[i] = 0; }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int item0 = tile * SAMPLE_ITEMS;
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

        // phase 2: the tile's slots, U neighbour reads in flight per thread
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
                    const int32_t d = dst[u];
                    int32_t h = -1;
                    if (d >= 0 && d < n_nodes) {                                   // Kernels.cu:411
                        h = dedup_claim(dd, d, cand_base | (int32_t)(slot0 + s), keep);
                        if (h < 0) st->status = LGN_E_CAPACITY;
                        else if (PRESC) atomicAdd(&s_cnt[il[u]], 1);
                    }
                    LGN_ASSERT(s < SAMPLE_ITEMS * f);
                    s_val[s] = h;
                }
            }
        }
        __syncthreads();

        // phase 3: compact the valid draws in slot order into the tile's own region (no cross-tile prefix needed)
        const int per_warp = (total + SAMPLE_WARPS - 1) / SAMPLE_WARPS;
        const int w_lo = min(total, warp * per_warp), w_hi = min(total, w_lo + per_warp);
        const int w_end = w_lo + ((w_hi - w_lo + 31) & ~31);     // whole warp iterations: the ballots need every lane
        int cnt = 0;
        for (int s = w_lo + lane; s < w_end; s += 32)
            cnt += __popc(__ballot_sync(0xffffffffu, s < w_hi && s_val[s] >= 0));
        if (lane == 0) s_wcnt[warp] = cnt;
        __syncthreads();
        int pos = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < SAMPLE_WARPS; w++) {
            if (w < warp) pos += s_wcnt[w];
            tile_total += s_wcnt[w];
        }
        for (int s = w_lo + lane; s < w_end; s += 32) {
            const int32_t h = s < w_hi ? s_val[s] : -1;
            const uint32_t m = __ballot_sync(0xffffffffu, h >= 0);
            if (h >= 0) {
                const int j = pos + __popc(m & ((1u << lane) - 1u));
                LGN_ASSERT(j < total && slot0 + j < st->dbg_max_slots);
                draw_h[slot0 + j] = h;
                draw_s[slot0 + j] = (uint16_t)s;
            }
            pos += __popc(m);
        }
        if (t == 0) tile_n[tile] = tile_total;
        if (PRESC) {   // one global atomic per frontier item instead of one per edge (Kernels.cu:525)
            if (t < n_items && s_cnt[t] > 0) atomicAdd(&topo_hot[s_src[t]], (uint32_t)s_cnt[t]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ mark
// pass 1 over the hop's valid draws: who won?  One WARP per tile (no block barriers), U entries per lane in flight.
// Leaves one (edges, new nodes) count per tile and their sums per group of 64 tiles ("supertile"), so that pass 2 can
// form any tile's exclusive prefix from ~(tiles/64 + 64) values instead of a device-wide scan.
constexpr int SUPER = 64;

__global__ void __launch_bounds__(SCAN_THREADS) k_mark(const int32_t* __restrict__ draw_h, const uint16_t* __restrict__ draw_s,
                                                       int32_t* __restrict__ draw_v, int32_t* __restrict__ draw_key,
                                                       const int32_t* __restrict__ tile_n, int32_t* __restrict__ tile_new,
                                                       int32_t* __restrict__ super_e, int32_t* __restrict__ super_n, const Dedup dd,
                                                       const BatchState* __restrict__ st, int hop, int f)
{
    constexpr int U = 8;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * SCAN_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * SCAN_THREADS) >> 5;
    const int F = st->hop[hop].n_items;
    const int n_tiles = (F + SAMPLE_ITEMS - 1) / SAMPLE_ITEMS;
    const int TS = SAMPLE_ITEMS * f;
    const unsigned long long keep = policy_evict_last();
    for (int tile = warp; tile < n_tiles; tile += n_warps) {
        const int n = tile_n[tile];
        const long long base = (long long)tile * TS;
        int wins = 0;
        for (int j0 = 0; j0 < n; j0 += 32 * U) {
            int32_t h[U], pv[U], key[U];
            int s[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = j0 + 32 * u + lane;
                h[u] = j < n ? draw_h[base + j] : -1;
                s[u] = j < n ? (int)draw_s[base + j] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) { key[u] = -1; pv[u] = 0; if (h[u] >= 0) pv[u] = dedup_payload_key(dd, h[u], key[u], keep); }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = j0 + 32 * u + lane;
                if (h[u] >= 0) {
                    LGN_ASSERT(base + j < st->dbg_max_slots && (dd.bits == 0 || (uint32_t)h[u] < (1u << dd.bits)));
                    draw_v[base + j] = pv[u];
                    if (dd.bits) draw_key[base + j] = key[u];
                    wins += pv[u] == (CAND | (int32_t)(base + s[u])) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wins += __shfl_xor_sync(0xffffffffu, wins, o);
        if (lane == 0) {
            tile_new[tile] = wins;
            if (n) atomicAdd(&super_e[tile / SUPER], n);
            if (wins) atomicAdd(&super_n[tile / SUPER], wins);
        }
    }
}

// ------------------------------------------------------------------ assign
// pass 2, one warp per tile: edge index = tile prefix + position, winners numbered by ballot scans; both COO index
// arrays written.  The warp that owns the LAST tile advances the counters (update_counter(op 2/4), Kernels.cu:128-149).
__device__ __forceinline__ int warp_sum(int x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_assign(
    const int32_t* __restrict__ draw_h, const uint16_t* __restrict__ draw_s, const int32_t* __restrict__ draw_v,
    const int32_t* __restrict__ draw_key, const int32_t* __restrict__ tile_n, const int32_t* __restrict__ tile_new,
    const int32_t* __restrict__ super_e, const int32_t* __restrict__ super_n, int32_t* __restrict__ ids,
    int32_t* __restrict__ agg_src_ids, int32_t* __restrict__ agg_dst_ids, int32_t* __restrict__ agg_src_off,
    int32_t* __restrict__ agg_dst_off, const Dedup dd, const int32_t* __restrict__ seed_h, int32_t* __restrict__ nc,
    int32_t* __restrict__ ec, BatchState* __restrict__ st, int hop, int f, long long capacity)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * SCAN_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * SCAN_THREADS) >> 5;
    const HopState hs = st->hop[hop];
    const int F = hs.n_items;
    const int n_tiles = (F + SAMPLE_ITEMS - 1) / SAMPLE_ITEMS;
    const int TS = SAMPLE_ITEMS * f;
    const int32_t gen_base = st->gen_base;
    const int32_t* frontier = hop == 0 ? ids : agg_src_ids + hs.item_base;
    const unsigned long long keep = policy_evict_last();
    // an empty frontier has no last tile: warp 0 advances the counters
    const int tile_end = n_tiles > 0 ? n_tiles : 1;
    for (int tile = warp; tile < tile_end; tile += n_warps) {
        // exclusive prefix of (edges, new nodes) before this tile: whole supertiles, then the tiles of its own supertile
        int pe = 0, pn = 0;
        const int sup = tile / SUPER;
        for (int i = lane; i < sup; i += 32) { pe += super_e[i]; pn += super_n[i]; }
        for (int i = sup * SUPER + lane; i < tile; i += 32) { pe += tile_n[i]; pn += tile_new[i]; }
        pe = warp_sum(pe); pn = warp_sum(pn);
        const int n = n_tiles > 0 ? tile_n[tile] : 0;
        const long long base = (long long)tile * TS;
        const long long e0 = (long long)hs.edge_base + pe;
        long long p0 = (long long)hs.node_base + pn;
        for (int j0 = 0; j0 < n; j0 += 32 * U) {
            int32_t h[U], pv[U], key[U], src[U], dst_off[U];
            int s[U], rank[U];
            bool win[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int j = j0 + 32 * u + lane;
                h[u] = -1; pv[u] = 0; s[u] = 0; key[u] = -1;
                if (j < n) {
                    h[u] = draw_h[base + j]; pv[u] = draw_v[base + j]; s[u] = draw_s[base + j];
                    key[u] = dd.bits ? draw_key[base + j] : h[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {   // frontier node of the draw and its local index
                src[u] = -1; dst_off[u] = 0;
                if (h[u] >= 0) {
                    const int item = tile * SAMPLE_ITEMS + s[u] / f;
                    src[u] = frontier[item];
                    // a seed's index sits in its dedup entry; a later hop's item is the previous hop's edge, already relabelled
                    // (construct_graph, Kernels.cu:458-461) unless its winner was numbered by another thread: then the entry
                    // -(handle+2) points at the dedup entry that holds the index by now
                    if (hop == 0) dst_off[u] = dd.bits ? seed_h[item] : src[u];      // handle of the seed (hash: -1 if the table was full)
                    else dst_off[u] = agg_src_off[hs.item_base + item];
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (h[u] >= 0) {
                    const int item = tile * SAMPLE_ITEMS + s[u] / f;
                    if (hop == 0) dst_off[u] = dst_off[u] >= 0 ? dedup_payload(dd, dst_off[u], keep) : item;
                    else if (dst_off[u] < 0) dst_off[u] = dedup_payload(dd, -2 - dst_off[u], keep);
                }
                win[u] = h[u] >= 0 && pv[u] == (CAND | (int32_t)(base + s[u]));
                const uint32_t m = __ballot_sync(0xffffffffu, win[u]);
                rank[u] = (int)(p0 - ((long long)hs.node_base + pn)) + __popc(m & ((1u << lane) - 1u));
                p0 += __popc(m);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (h[u] < 0) continue;
                const int j = j0 + 32 * u + lane;
                int32_t src_off;
                if (win[u]) {                       // Kernels.cu:418-438
                    const long long pos = (long long)hs.node_base + pn + rank[u];
                    LGN_ASSERT(pos >= hs.node_base && pos <= st->dbg_capacity && key[u] >= 0);
                    if (pos < capacity) {
                        ids[pos] = key[u];
                        dedup_publish(dd, h[u], key[u], gen_base | (int32_t)pos, keep);
                    }
                    src_off = (int32_t)pos;
                } else {
                    src_off = pv[u] < CAND ? pv[u] : -2 - h[u];   // winner of this hop numbered elsewhere: read it through the entry later
                }
                const long long e = e0 + j;
                LGN_ASSERT(e >= hs.edge_base && e <= st->dbg_capacity && dst_off[u] >= 0 && (src_off >= 0 || -2 - src_off == h[u]));
                if (e < capacity) {                    // Kernels.cu:423-424, 441-445
                    agg_src_ids[e] = key[u];
                    agg_dst_ids[e] = src[u];
                    agg_src_off[e] = src_off;
                    agg_dst_off[e] = dst_off[u];
                }
            }
        }
        if (tile == tile_end - 1 && lane == 0) {
            const int te = pe + n, tn = (int)(p0 - (long long)hs.node_base);
            HopState nx;
            nx.n_items = te; nx.item_base = hs.edge_base; nx.node_base = hs.node_base + tn; nx.edge_base = hs.edge_base + te;
            st->hop[hop + 1] = nx;
            st->tot_items += (unsigned long long)hs.n_items;
            st->tot_edges += (unsigned long long)te;
            if ((long long)nx.node_base > capacity || (long long)nx.edge_base > capacity) st->status = LGN_E_CAPACITY;
            nc[0] = nx.node_base; nc[1] = 0; nc[2] = te;
            nc[5 + 2 * hop] = hs.node_base; nc[6 + 2 * hop] = tn;
            if (7 + 2 * hop < 16) nc[7 + 2 * hop] = nx.node_base;
            ec[0] = nx.edge_base; ec[1] = 0; ec[2] = hs.edge_base; ec[3 + hop] = nx.edge_base;
        }
    }
}

// ------------------------------------------------------------------ batch end
template <bool PRESC>
__global__ void __launch_bounds__(256) k_batch_end(const int32_t* __restrict__ ids, const Dedup dd,
                                                   BatchState* __restrict__ st, int n_hops,
                                                   uint32_t* __restrict__ node_hot, long long n_nodes,
                                                   int32_t* __restrict__ agg_src_off)
{
    // in-hop duplicates numbered by another thread: the index is in the node's dedup entry (construct_graph's second
    // lookup, Kernels.cu:458).  Four edges per thread and iteration: one 16-byte load, the (rare) probes of all four in
    // flight together, only patched entries written back.
    const int n_edges = st->hop[n_hops].edge_base;
    const unsigned long long keep = policy_evict_last();
    const int n_quads = (n_edges + 3) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
        const int e = q << 2;
        int32_t so[4];
        if (e + 3 < n_edges) {
            const int4 v = *reinterpret_cast<const int4*>(agg_src_off + e);
            so[0] = v.x; so[1] = v.y; so[2] = v.z; so[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) so[k] = e + k < n_edges ? agg_src_off[e + k] : 0;
        }
        int32_t fix[4];
#pragma unroll
        for (int k = 0; k < 4; k++) fix[k] = so[k] < 0 ? dedup_payload(dd, -2 - so[k], keep) : so[k];
#pragma unroll
        for (int k = 0; k < 4; k++) if (so[k] < 0) { LGN_ASSERT(e + k < n_edges && fix[k] >= 0 && fix[k] < CAND); agg_src_off[e + k] = fix[k]; }
    }
    if (PRESC) {
        const int total = st->hop[n_hops].node_base;   // nc[9] for two hops
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const int32_t id = ids[i];
            if (id >= 0 && id < n_nodes) atomicAdd(&node_hot[id], 1u);   // HotnessMeasure: ids of a batch are unique -> no contention
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && total > st->max_ids) st->max_ids = total;
    }
}

// ------------------------------------------------------------------ launchers
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

void launch_batch_begin(lgn_ctx* c, cudaStream_t s, const int32_t* ids, const int32_t* labels, int32_t src_off,
                        int32_t count, uint32_t step)
{
    Pipe& p = c->pipe[c->cur_pipe];
    step += c->rng_step_offset;
    // generations run 62, 61, .., 0; before the first batch of the next cycle the dedup structure is reset (common.cuh)
    const int gen = N_GEN - 1 - (int)(p.batch_seq % N_GEN);
    if (p.batch_seq % N_GEN == 0 && p.batch_seq > 0) reset_dedup(c, p, s);
    p.batch_seq++;
    int blocks = cdiv(count, 256);
    if (blocks < 1) blocks = 1;
    k_batch_begin<<<blocks, 256, 0, s>>>(ids + src_off, labels ? labels + src_off : nullptr, count, p.ids, p.labels,
                                         p.dedup, p.seed_h, c->cfg.n_nodes, p.nc, p.ec, p.state, step, c->rng_epoch,
                                         (int32_t)(gen << GEN_SHIFT));
}

void launch_sample_hop(lgn_ctx* c, cudaStream_t s, int hop, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    long long fmax = c->cfg.batch_size;
    for (int h = 0; h < hop; h++) fmax *= c->cfg.fanout[h];
    const int f = c->cfg.fanout[hop];
    const int tiles_max = cdiv(fmax, SAMPLE_ITEMS);
    int sblocks = tiles_max;
    if (sblocks > c->n_sm * c->sample_ctas_per_sm) sblocks = c->n_sm * c->sample_ctas_per_sm;
    const size_t smem = (size_t)SAMPLE_ITEMS * f * sizeof(int32_t);     // <= 32 KB (f <= 256)
#define LGN_SAMPLE(R, P)                                                                                                  \
    k_sample<R, P><<<sblocks, SAMPLE_THREADS, smem, s>>>(c->topo, p.ids, p.agg_src_ids, p.draw_h, p.draw_s, p.tile_n, p.super_e, p.super_n, p.dedup, \
                                                         p.state, hop, f, c->cfg.rng_seed, c->topo_hotness, c->cfg.n_nodes)
    if (c->cfg.rng_mode == LGN_RNG_MINSTD) { if (presc) LGN_SAMPLE(LGN_RNG_MINSTD, true); else LGN_SAMPLE(LGN_RNG_MINSTD, false); }
    else { if (presc) LGN_SAMPLE(LGN_RNG_PHILOX, true); else LGN_SAMPLE(LGN_RNG_PHILOX, false); }
#undef LGN_SAMPLE
    int rblocks = cdiv(tiles_max, SCAN_THREADS / 32);      // one warp per tile
    if (rblocks > c->n_sm * c->resolve_ctas_per_sm) rblocks = c->n_sm * c->resolve_ctas_per_sm;
    if (rblocks < 1) rblocks = 1;
    k_mark<<<rblocks, SCAN_THREADS, 0, s>>>(p.draw_h, p.draw_s, p.draw_v, p.draw_key, p.tile_n, p.tile_new, p.super_e, p.super_n, p.dedup,
                                            p.state, hop, f);
    k_assign<<<rblocks, SCAN_THREADS, 0, s>>>(p.draw_h, p.draw_s, p.draw_v, p.draw_key, p.tile_n, p.tile_new, p.super_e, p.super_n, p.ids,
                                              p.agg_src_ids, p.agg_dst_ids, p.agg_src_off, p.agg_dst_off, p.dedup, p.seed_h, p.nc, p.ec,
                                              p.state, hop, f, c->capacity);
}

void launch_batch_end(lgn_ctx* c, cudaStream_t s, bool presc)
{
    Pipe& p = c->pipe[c->cur_pipe];
    int blocks = cdiv(c->capacity, 256);
    if (blocks > c->n_sm * c->end_ctas_per_sm) blocks = c->n_sm * c->end_ctas_per_sm;
    if (presc) k_batch_end<true><<<blocks, 256, 0, s>>>(p.ids, p.dedup, p.state, c->cfg.n_hops, c->node_hotness, c->cfg.n_nodes, p.agg_src_off);
    else k_batch_end<false><<<blocks, 256, 0, s>>>(p.ids, p.dedup, p.state, c->cfg.n_hops, nullptr, c->cfg.n_nodes, p.agg_src_off);
}

int sample_items_per_tile() { return SAMPLE_ITEMS; }

void sampler_set_carveout(int pct)
{
    cudaFuncSetAttribute(k_batch_begin, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_sample<LGN_RNG_MINSTD, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_sample<LGN_RNG_MINSTD, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_sample<LGN_RNG_PHILOX, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_sample<LGN_RNG_PHILOX, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_mark, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_assign, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_batch_end<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_batch_end<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace lgn
