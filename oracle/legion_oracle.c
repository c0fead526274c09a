/*
 * legion_oracle.c -- CPU restatement of Legion's mini-batch hot path.
 * TEST INFRASTRUCTURE ONLY (see legion_oracle.h).  Plain C11 (+ optional
 * OpenMP for the cpu_baseline leg; the threaded paths produce the same bytes
 * as the scalar ones, tests/test_oracle.py checks that).
 */
#include "legion_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ RNG */

#define MINSTD_A 48271ull
#define MINSTD_M 2147483647ull

/* thrust/random/detail/linear_congruential_engine_discard.h: square-and-multiply
 * with 64-bit intermediates; engine seeded with 1 so the state after discard(z)
 * is a^z, and the draw that follows returns a^(z+1). */
uint32_t lgo_minstd_pow(uint64_t e)
{
    uint64_t base = MINSTD_A, acc = 1;
    while (e) {
        if (e & 1) acc = (acc * base) % MINSTD_M;
        e >>= 1;
        base = (base * base) % MINSTD_M;
    }
    return (uint32_t)acc;
}

/* Kernels.cu:402-405 + thrust uniform_int_distribution.inl /
 * uniform_real_distribution.inl: u = double(x - min)/(1.0 + double(max-min)),
 * result = int(u * double((deg-1)+1) + 0.0); min = 1, max = 2^31-2. */
int32_t lgo_minstd_pick(uint64_t idx, int32_t deg)
{
    uint32_t x = lgo_minstd_pow(idx + 1);
    double u = (double)(x - 1u);
    u /= (1.0 + (double)(2147483646u - 1u));
    double hi = (double)(deg - 1) + 1.0;
    return (int32_t)(u * (hi - 0.0) + 0.0);
}

static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo)
{
    uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

void lgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo32(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int32_t lgo_philox_pick(uint64_t idx, uint32_t epoch, uint32_t hop, uint32_t step, uint64_t seed, int32_t deg)
{
    /* counter = (slot, epoch, hop, step): the slot index of a hop is < 2^30 (buffer capacity), so the word that
     * used to carry its (always zero) high half carries the epoch -- epoch 0 is the round-1 stream */
    uint32_t ctr[4] = { (uint32_t)idx, epoch + (uint32_t)(idx >> 32), hop, step };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    lgo_philox4x32_10(ctr, key, out);
    return (int32_t)(((uint64_t)out[0] * (uint32_t)deg) >> 32);
}

/* ---------------------------------------------------------- batch gen */

int32_t lgo_batch_generate(const int32_t* all_ids, const int32_t* all_labels,
                           int32_t total_cap, int32_t batch_size, int32_t counter,
                           int32_t* out_ids, int32_t* out_labels)
{
    /* Kernels.cu:224 */
    int32_t size = ((batch_size * (counter + 1)) >= total_cap) ? (total_cap - batch_size * counter)
                                                               : batch_size;
    if (size < 0) size = 0;
    /* Kernels.cu:227 passes `size` as the kernel's batch_size, so the kernel's
     * own index is size*counter+idx (Kernels.cu:81-93). */
    for (int32_t idx = 0; idx < size; idx++) {
        int64_t g = (int64_t)size * counter + idx;
        if (g >= total_cap) {
            out_ids[idx] = -1;
            if (out_labels) out_labels[idx] = -1;
        } else {
            out_ids[idx] = all_ids[g % total_cap];
            if (out_labels) out_labels[idx] = all_labels[g % total_cap];
        }
    }
    return size;
}

/* ------------------------------------------------------------ sampling */

static inline int32_t draw(const lgo_sample_args* a, uint32_t hop, int32_t f, int64_t idx,
                           int32_t k, int32_t deg)
{
    if (a->rng_mode == LGO_RNG_MINSTD) return lgo_minstd_pick((uint64_t)idx, deg);
    /* philox mode: exact neighbourhood when the fanout covers the degree */
    if (deg <= f) return k;
    return lgo_philox_pick((uint64_t)idx, a->epoch, hop, a->step, a->rng_seed, deg);
}

int lgo_sample_batch(lgo_sample_args* a)
{
    int32_t* nc = a->nc;
    int32_t* ec = a->ec;
    int32_t* pos = a->position_map;
    const int32_t B = a->n_seeds;
    if (a->n_hops < 0 || a->n_hops > 5 || B < 0 || B > a->capacity) return LGO_E_ARG;

    /* batch_generator + update_counter(op 0): Kernels.cu:68-96, 118-127 */
    memset(nc, 0, 16 * sizeof(int32_t));
    memset(ec, 0, 16 * sizeof(int32_t));
    for (int32_t i = 0; i < B; i++) {
        int32_t id = a->sampled_ids[i];
        if (id >= 0 && pos[id] < 0) pos[id] = i; /* canonical: lowest index wins */
    }
    nc[0] = B; nc[1] = 0; nc[2] = B; nc[3] = 0; nc[4] = B;

    int rc = LGO_OK;
    for (int32_t h = 0; h < a->n_hops && rc == LGO_OK; h++) {
        const int32_t f = a->fanout[h];
        /* Kernels.cu:368-374: hop 1 expands the seeds, later hops expand the
         * previous hop's edge list (duplicates included). */
        const int32_t* frontier = (h == 0) ? a->sampled_ids : a->agg_src_ids + ec[2];
        const int64_t F = nc[2];
        const int64_t slots = F * (int64_t)f;
        const int64_t nbase = nc[0], ebase = ec[0];
        int64_t n_new = 0, n_e = 0;

        int32_t* picked = NULL;
#ifdef _OPENMP
        if (a->n_threads > 1 && slots > 0) {
            picked = (int32_t*)malloc((size_t)slots * sizeof(int32_t));
#pragma omp parallel for schedule(dynamic, 4096) num_threads(a->n_threads)
            for (int64_t idx = 0; idx < slots; idx++) {
                int64_t i = idx / f; int32_t k = (int32_t)(idx % f);
                int32_t src = frontier[i];
                int32_t dst = -1;
                if (src >= 0) {
                    int64_t start = a->indptr[src];
                    int32_t deg = (int32_t)(a->indptr[src + 1] - start);
                    if (k < deg) dst = a->indices[start + draw(a, (uint32_t)h, f, idx, k, deg)];
                }
                picked[idx] = dst;
            }
        }
#endif
        for (int64_t idx = 0; idx < slots; idx++) {
            int64_t i = idx / f; int32_t k = (int32_t)(idx % f);
            int32_t src = frontier[i];
            if (src < 0) continue;                          /* Kernels.cu:385 */
            int32_t dst;
            if (picked) {
                dst = picked[idx];
            } else {
                int64_t start = a->indptr[src];
                int32_t deg = (int32_t)(a->indptr[src + 1] - start);
                if (k >= deg) continue;                     /* Kernels.cu:399 */
                dst = a->indices[start + draw(a, (uint32_t)h, f, idx, k, deg)];
            }
            if (dst < 0) continue;                          /* Kernels.cu:411 */
            if (a->topo_hotness) a->topo_hotness[src] += 1; /* Kernels.cu:525 */
            if (pos[dst] < 0) {                             /* Kernels.cu:415-421 */
                if (nbase + n_new >= a->capacity) { rc = LGO_E_CAPACITY; break; }
                pos[dst] = (int32_t)(nbase + n_new);
                a->sampled_ids[nbase + n_new] = dst;
                n_new++;
            }
            if (ebase + n_e >= a->capacity) { rc = LGO_E_CAPACITY; break; }
            int64_t e = ebase + n_e++;
            a->agg_src_ids[e] = dst;                        /* Kernels.cu:423-424 */
            a->agg_dst_ids[e] = src;
            a->agg_src_off[e] = pos[dst];                   /* construct_graph, :450-463 */
            a->agg_dst_off[e] = pos[src];
        }
        free(picked);

        /* update_counter(op 2 / op 4), Kernels.cu:128-149, generalised to hop h:
         * segment h+1 = (nc[5+2h], nc[6+2h]), running total in nc[7+2h],
         * cumulative edges in ec[3+h]. */
        nc[0] += (int32_t)n_new;
        nc[5 + 2 * h] = nc[3 + 2 * h] + nc[4 + 2 * h];
        nc[6 + 2 * h] = (int32_t)n_new;
        if (7 + 2 * h < 16) nc[7 + 2 * h] = nc[5 + 2 * h] + nc[6 + 2 * h];
        nc[1] = 0;
        nc[2] = (int32_t)n_e;
        ec[3 + h] = (h == 0 ? 0 : ec[3 + h - 1]) + (int32_t)n_e;
        ec[2] = ec[0];
        ec[0] += (int32_t)n_e;
        ec[1] = 0;
    }

    /* HotnessMeasure (GPUCache.cu:227-235): one count per unique id of the batch */
    const int32_t total = nc[0];
    if (a->node_hotness)
        for (int32_t i = 0; i < total; i++)
            if (a->sampled_ids[i] >= 0) a->node_hotness[a->sampled_ids[i]] += 1;
    /* restore scratch (the reference's ClearPosMap + per-batch bitmap memset) */
    for (int32_t i = 0; i < total; i++)
        if (a->sampled_ids[i] >= 0) pos[a->sampled_ids[i]] = -1;
    return rc;
}

void lgo_draw_hop(const int64_t* indptr, const int32_t* indices, const int32_t* frontier, int64_t n_items,
                  int32_t f, int32_t rng_mode, uint64_t rng_seed, uint32_t hop, uint32_t step, uint32_t epoch, int32_t* out_dst)
{
    lgo_sample_args a;
    memset(&a, 0, sizeof(a));
    a.rng_mode = rng_mode; a.rng_seed = rng_seed; a.step = step; a.epoch = epoch;
    for (int64_t idx = 0; idx < n_items * f; idx++) {
        int64_t i = idx / f; int32_t k = (int32_t)(idx % f);
        int32_t src = frontier[i], dst = -1;
        if (src >= 0) {
            int64_t start = indptr[src];
            int32_t deg = (int32_t)(indptr[src + 1] - start);
            if (k < deg) dst = indices[start + draw(&a, hop, f, idx, k, deg)];
        }
        out_dst[idx] = dst;
    }
}

/* ------------------------------------------------------------- planner */

typedef struct { uint32_t c; int32_t id; } hot_pair;

static int hot_cmp(const void* x, const void* y)
{
    const hot_pair* a = (const hot_pair*)x; const hot_pair* b = (const hot_pair*)y;
    if (a->c != b->c) return a->c > b->c ? -1 : 1;
    return a->id < b->id ? -1 : (a->id > b->id);
}

/* GPUCache.cu:630-631: iota payload + sort_by_key(greater) through cub's stable
 * radix sort == (count desc, id asc). */
void lgo_hot_order(const uint32_t* counts, int64_t n, int32_t* order)
{
    hot_pair* p = (hot_pair*)malloc((size_t)n * sizeof(hot_pair));
    for (int64_t i = 0; i < n; i++) { p[i].c = counts[i]; p[i].id = (int32_t)i; }
    qsort(p, (size_t)n, sizeof(hot_pair), hot_cmp);
    for (int64_t i = 0; i < n; i++) order[i] = p[i].id;
    free(p);
}

void lgo_place(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t* slot_of)
{
    for (int64_t i = 0; i < n; i++) slot_of[i] = -1;
    int64_t lim = cap * kg < n ? cap * kg : n;
    for (int64_t i = 0; i < lim; i++)
        slot_of[order[i]] = (int32_t)((i % kg) * cap + i / kg);   /* GPUCache.cu:106 */
}

void lgo_fill_feature_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                            const float* features, int32_t dim, float* shard)
{
    for (int64_t r = 0; r < cap; r++) {
        int64_t rank = r * kg + j;                                /* GPUCache.cu:202 */
        if (rank >= n) continue;
        memcpy(shard + r * dim, features + (int64_t)order[rank] * dim, (size_t)dim * sizeof(float));
    }
}

void lgo_place_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int64_t n_repl, int32_t my_part, int32_t* slot_of)
{
    int64_t lim = n_repl + (cap - n_repl) * kg;
    for (int64_t i = 0; i < n; i++) {
        int32_t slot = -1;
        if (i < n_repl) slot = (int32_t)(my_part * cap + i);
        else if (i < lim) { int64_t k = i - n_repl; slot = (int32_t)((k % kg) * cap + n_repl + k / kg); }
        slot_of[order[i]] = slot;
    }
}

void lgo_fill_feature_shard_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j, int64_t n_repl,
                                   const float* features, int32_t dim, float* shard)
{
    for (int64_t r = 0; r < cap; r++) {
        int64_t rank = r < n_repl ? r : n_repl + (r - n_repl) * kg + j;
        if (rank >= n) continue;
        memcpy(shard + r * dim, features + (int64_t)order[rank] * dim, (size_t)dim * sizeof(float));
    }
}

int64_t lgo_fill_topo_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                            const int64_t* indptr, const int32_t* indices,
                            int64_t* indptr_out, int32_t* indices_out)
{
    int64_t acc = 0;
    indptr_out[0] = 0;
    for (int64_t t = 0; t < cap; t++) {
        int64_t rank = t * kg + j;       /* GPU_Memory_Graph_Storage.cu:16,26 */
        int64_t cnt = 0;
        if (rank < n) {
            int32_t id = order[rank];
            cnt = indptr[id + 1] - indptr[id];
            if (indices_out) memcpy(indices_out + acc, indices + indptr[id], (size_t)cnt * sizeof(int32_t));
        }
        acc += cnt;
        indptr_out[t + 1] = acc;
    }
    return acc;
}

void lgo_cost_model(const uint64_t* af, const uint64_t* at, const int32_t* qt,
                    const int64_t* indptr, int64_t n, int32_t dim, int64_t cache_memory,
                    int32_t kg, uint64_t topo_trans, const int32_t* max_ids, int32_t train_step,
                    int32_t* node_capacity, int32_t* edge_capacity, int32_t* best_step)
{
    /* GPUCache.cu:672-679 */
    const int max_payload = 64;
    int64_t memory_step = (int64_t)((double)(cache_memory * kg) * 0.01);
    if (memory_step < 1) memory_step = 1;
    uint64_t feat_trans = 0;
    for (int j = 0; j < kg; j++)
        feat_trans += (uint64_t)(((((int64_t)max_ids[j] * train_step) * dim) * (int64_t)sizeof(float)) / max_payload);

    uint64_t* node_prefix = (uint64_t*)malloc((size_t)n * 8);
    uint64_t* edge_prefix = (uint64_t*)malloc((size_t)n * 8);
    uint64_t* mem_prefix = (uint64_t*)malloc((size_t)n * 8);
    uint64_t s0 = 0, s1 = 0, s2 = 0;
    for (int64_t i = 0; i < n; i++) {
        s0 += af[i]; node_prefix[i] = s0;
        s1 += at[i]; edge_prefix[i] = s1;
        int32_t id = qt[i];
        s2 += 8u + 4u * (uint64_t)(indptr[id + 1] - indptr[id]);    /* GPUCache.cu:35-41 */
        mem_prefix[i] = s2;
    }

    int64_t total_mem = cache_memory * kg;
    int64_t steps = (total_mem - 1) / memory_step + 1;
    float* t_topo = (float*)calloc((size_t)steps + 1, sizeof(float));
    float* t_feat = (float*)calloc((size_t)steps + 1, sizeof(float));
    float* c_topo = (float*)calloc((size_t)steps + 1, sizeof(float));
    float* c_feat = (float*)calloc((size_t)steps + 1, sizeof(float));
    float* t_total = (float*)calloc((size_t)steps + 1, sizeof(float));
    int64_t cs = 0;
    for (int64_t cur = 0; cur < total_mem; cur += memory_step) {           /* :723-753 */
        int32_t nf, nt;
        if ((uint64_t)cur > (uint64_t)n * dim * sizeof(float)) nf = (int32_t)n;
        else nf = (int32_t)((uint64_t)(cs + 1) * ((uint64_t)memory_step / (dim * sizeof(float))));
        if ((uint64_t)cur > mem_prefix[n - 1]) nt = (int32_t)n;
        else {  /* std::lower_bound */
            int64_t lo = 0, hi = n;
            while (lo < hi) { int64_t mid = (lo + hi) / 2; if (mem_prefix[mid] < (uint64_t)cur) lo = mid + 1; else hi = mid; }
            nt = (int32_t)lo;
        }
        if (nt < n) {
            /* the reference reads prefix[-1] when nt == 0 (step 0 only, value never used) */
            uint64_t ep = nt > 0 ? edge_prefix[nt - 1] : 0;
            t_topo[cs] = (float)((double)topo_trans * 1.0 / (double)edge_prefix[n - 1] * (double)ep);
            c_topo[cs] = (float)(nt / kg);
        }
        if (nf < n) {
            uint64_t np = nf > 0 ? node_prefix[nf - 1] : 0;
            t_feat[cs] = (float)((double)feat_trans * 1.0 / (double)node_prefix[n - 1] * (double)np);
            c_feat[cs] = (float)(nf / kg);
        }
        cs++;
    }
    for (int64_t s = 1; s < steps; s++) t_total[s] = t_topo[s] + t_feat[steps - 1 - s];
    int64_t best = 0;
    for (int64_t s = 1; s <= steps; s++) if (t_total[s] > t_total[best]) best = s;  /* max_element: first max */
    *node_capacity = (int32_t)(c_feat[steps - 1 - best] + 1);
    *edge_capacity = (int32_t)(c_topo[best] + 1);
    if (best_step) *best_step = (int32_t)best;
    free(node_prefix); free(edge_prefix); free(mem_prefix);
    free(t_topo); free(t_feat); free(c_topo); free(c_feat); free(t_total);
}

/* -------------------------------------------------------------- gather */

void lgo_gather(const int32_t* sampled_ids, int32_t off, int32_t cnt,
                const int32_t* slot_of, int64_t cap, const float* const* shards,
                const float* host_features, int64_t n_nodes, int32_t dim, float* out,
                int64_t* tier_rows, int32_t n_shards, int32_t n_threads)
{
    (void)n_threads;
    if (tier_rows) memset(tier_rows, 0, (size_t)(n_shards + 1) * sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) if (n_threads > 1 && !tier_rows)
#endif
    for (int32_t r = 0; r < cnt; r++) {
        int32_t id = sampled_ids[off + r];
        float* dst = out + (int64_t)(off + r) * dim;
        int32_t g = -1;
        if (id >= 0 && slot_of) g = slot_of[id % n_nodes];          /* Find_Kernel semantics, GPUCache.cu:170-176 */
        if (g < 0) {                                                /* Kernels.cu:692-696 */
            if (id >= 0) {
                memcpy(dst, host_features + (int64_t)(id % n_nodes) * dim, (size_t)dim * sizeof(float));
                if (tier_rows) tier_rows[n_shards]++;
            }
        } else {                                                    /* Kernels.cu:697-699 */
            int64_t d = g / cap, row = g % cap;
            memcpy(dst, shards[d] + row * dim, (size_t)dim * sizeof(float));
            if (tier_rows) tier_rows[d]++;
        }
    }
}

/* ------------------------------------------------------ step arithmetic */

void lgo_coordinate(const int32_t* n_train, const int32_t* n_valid, const int32_t* n_test,
                    int32_t parts, int32_t batch, int32_t epochs, lgo_steps* out)
{
    int32_t min_train = 1000000000, max_valid = 0, max_test = 0;      /* CUDA_IPC_Service.cu:71-112 */
    for (int i = 0; i < parts; i++) {
        if (n_train[i] < min_train) min_train = n_train[i];
        if (n_valid[i] > max_valid) max_valid = n_valid[i];
        if (n_test[i] > max_test) max_test = n_test[i];
    }
    out->train_step = (min_train - 1) / batch;
    out->valid_step = (max_valid - 1) / 512 + 1;
    out->test_step = (max_test - 1) / 512 + 1;
    for (int i = 0; i < parts && i < 8; i++) {
        out->valid_batch[i] = (n_valid[i] - 1) / out->valid_step + 1;
        out->test_batch[i] = (n_test[i] - 1) / out->test_step + 1;
    }
    out->max_step = (out->train_step + out->valid_step) * epochs + out->test_step;
}

int32_t lgo_mode_of_step(const lgo_steps* s, int32_t epochs, int32_t g)
{
    if (g < (s->train_step + s->valid_step) * epochs)                 /* :246-259 */
        return (g % (s->train_step + s->valid_step)) < s->train_step ? 0 : 1;
    return 2;
}

int32_t lgo_local_batch_id(const lgo_steps* s, int32_t epochs, int32_t g)
{
    if (g < (s->train_step + s->valid_step) * epochs) {               /* :219-233 */
        int32_t e = g % (s->train_step + s->valid_step);
        return e < s->train_step ? e : e - s->train_step;
    }
    return (g - (s->train_step + s->valid_step) * epochs) % s->test_step;
}
