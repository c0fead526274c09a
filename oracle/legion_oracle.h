/*
 * legion_oracle.h -- CPU restatement of Legion's mini-batch hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call this.
 * The product path (legion-1_b200/csrc, include/legion_b200.h) never does.
 *
 * Every function cites the reference file:line (relative to the reference
 * checkout) whose behaviour it restates.  Where the reference's output order
 * depends on atomic arrival order (Kernels.cu:418-445) the oracle fixes the
 * canonical order "first occurrence in slot order" (slot idx = item*fanout+k),
 * which is one of the orders the reference itself can produce.
 *
 * Pins: thrust::minstd_rand closed form checked against the CUDA toolkit's own
 * thrust headers compiled for the host (oracle/pins/minstd_pin.cpp ->
 * tests/golden/minstd_pin.json), Philox4x32-10 against the Random123 known
 * answers and libcu++'s philox4x32 (oracle/pins/philox_pin.cpp), and the whole
 * sampling / gather path against the reference's own kernels recompiled for
 * sm_100a and run on the B200 box (oracle/ref_harness, tests/test_reference_ab.py).
 */
#ifndef LEGION_ORACLE_H
#define LEGION_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LGO_RNG_MINSTD = 0, LGO_RNG_PHILOX = 1 };
enum { LGO_OK = 0, LGO_E_CAPACITY = -2, LGO_E_ARG = -1 };

/* ---- RNG ------------------------------------------------------------- */
/* 48271^e mod (2^31-1): thrust::minstd_rand state after discard(e-1) from seed 1
 * (thrust/random/detail/linear_congruential_engine_discard.h). */
uint32_t lgo_minstd_pow(uint64_t e);
/* Kernels.cu:402-405: minstd_rand engine; engine.discard(idx);
 * uniform_int_distribution<>(0,deg-1)(engine). */
int32_t lgo_minstd_pick(uint64_t idx, int32_t deg);
/* Philox4x32-10 (Salmon et al., SC'11), one block. */
void lgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* counter = (slot, epoch, hop, step), key = (seed_lo, seed_hi); pick = (out[0] * deg) >> 32.
 * slot = item*f + k inside the hop (< 2^30), epoch/step = position of the mini-batch in the run
 * (stream keyed by seed, epoch, step, hop, slot). */
int32_t lgo_philox_pick(uint64_t idx, uint32_t epoch, uint32_t hop, uint32_t step, uint64_t seed, int32_t deg);

/* ---- batch generation (Kernels.cu:68-96, 163-232) --------------------- */
/* Returns the actual number of seeds of step `counter` (Kernels.cu:224) and
 * fills ids/labels.  Keeps the reference's index arithmetic, including its use
 * of the clamped size as the stride for the last partial batch. */
int32_t lgo_batch_generate(const int32_t* all_ids, const int32_t* all_labels,
                           int32_t total_cap, int32_t batch_size, int32_t counter,
                           int32_t* out_ids, int32_t* out_labels);

/* ---- k-hop sampling (Kernels.cu:112-150, 342-463, 468-564) ------------ */
typedef struct {
    /* graph */
    int64_t n_nodes;
    const int64_t* indptr;   /* int64[n_nodes+1] */
    const int32_t* indices;  /* int32[E] */
    /* sampling config */
    int32_t n_hops;
    const int32_t* fanout;   /* int32[n_hops] */
    int32_t rng_mode;        /* LGO_RNG_* */
    uint64_t rng_seed;       /* philox only */
    uint32_t step;           /* philox only: batch id inside the epoch */
    uint32_t epoch;          /* philox only: epoch (0 = first) */
    /* batch buffers, capacity entries each (labels not touched here) */
    int64_t capacity;
    int32_t n_seeds;         /* sampled_ids[0..n_seeds) already holds the seeds */
    int32_t* sampled_ids;
    int32_t* agg_src_ids;    /* raw id of the sampled neighbour (message source) */
    int32_t* agg_dst_ids;    /* raw id of the frontier node */
    int32_t* agg_src_off;    /* local index of neighbour */
    int32_t* agg_dst_off;    /* local index of frontier node */
    int32_t* nc;             /* int32[16] node_counter */
    int32_t* ec;             /* int32[16] edge_counter */
    /* scratch: int32[n_nodes], all -1 on entry, restored to -1 on exit */
    int32_t* position_map;
    /* optional presampling histograms (may be NULL) */
    uint32_t* topo_hotness;  /* += 1 per sampled edge out of src (Kernels.cu:525) */
    uint32_t* node_hotness;  /* += 1 per unique id of the batch (GPUCache.cu:227-235) */
    int32_t n_threads;       /* >1: OpenMP over slot draws (same output) */
} lgo_sample_args;

int lgo_sample_batch(lgo_sample_args* a);

/* the draw step alone for one hop over an explicit frontier (Kernels.cu:383-410):
 * out_dst[i*f+k] = sampled neighbour or -1.  Lets tests replay the reference's own (atomic-
 * arrival) frontier order, on which its hop>=2 draws depend through idx. */
void lgo_draw_hop(const int64_t* indptr, const int32_t* indices, const int32_t* frontier, int64_t n_items,
                  int32_t f, int32_t rng_mode, uint64_t rng_seed, uint32_t hop, uint32_t step, uint32_t epoch, int32_t* out_dst);

/* ---- cache planning (GPUCache.cu:578-659, 88-108, 200-205) ------------ */
/* order[i] = node of rank i under (count desc, id asc). */
void lgo_hot_order(const uint32_t* counts, int64_t n, int32_t* order);
/* slot_of[id] = (i%Kg)*cap + i/Kg for rank i < min(cap*Kg, n); -1 otherwise.
 * (InitPair, GPUCache.cu:103-108) */
void lgo_place(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t* slot_of);
/* shard j, row r <- features[order[r*Kg+j]] (FeatFillUp, GPUCache.cu:200-205).
 * Rows whose rank is >= n are left untouched. */
void lgo_fill_feature_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                            const float* features, int32_t dim, float* shard);
/* hybrid placement (extension of this repo, no reference counterpart): the n_repl hottest ranks are replicated
 * on every GPU (rows [0,n_repl)), the following (cap-n_repl)*kg ranks partitioned round-robin into rows [n_repl,cap). */
void lgo_place_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int64_t n_repl, int32_t my_part, int32_t* slot_of);
void lgo_fill_feature_shard_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j, int64_t n_repl,
                                   const float* features, int32_t dim, float* shard);
/* shard j topology CSR of nodes order[t*Kg+j] (GetNeighborCount/TopoFillUp,
 * GPU_Memory_Graph_Storage.cu:14-35,98-133).  indptr_out int64[cap+1];
 * indices_out may be NULL to only size it.  Returns number of indices. */
int64_t lgo_fill_topo_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                            const int64_t* indptr, const int32_t* indices,
                            int64_t* indptr_out, int32_t* indices_out);
/* CostModel (GPUCache.cu:661-767) for one clique.  af/at: hotness sorted
 * descending (feature / topology), qt: topology order.  Writes per-GPU
 * capacities. */
void lgo_cost_model(const uint64_t* af, const uint64_t* at, const int32_t* qt,
                    const int64_t* indptr, int64_t n, int32_t dim, int64_t cache_memory,
                    int32_t kg, uint64_t topo_trans, const int32_t* max_ids, int32_t train_step,
                    int32_t* node_capacity, int32_t* edge_capacity, int32_t* best_step);

/* ---- feature gather (Kernels.cu:662-702) ------------------------------ */
/* rows [off, off+cnt) of sampled_ids; slot_of may be NULL (all miss). */
void lgo_gather(const int32_t* sampled_ids, int32_t off, int32_t cnt,
                const int32_t* slot_of, int64_t cap, const float* const* shards,
                const float* host_features, int64_t n_nodes, int32_t dim, float* out,
                int64_t* tier_rows /* [0]=hit rows per shard.. may be NULL */, int32_t n_shards,
                int32_t n_threads);

/* ---- step arithmetic (CUDA_IPC_Service.cu:66-134, 219-259) ------------ */
typedef struct {
    int32_t train_step, valid_step, test_step, max_step;
    int32_t valid_batch[8], test_batch[8];
} lgo_steps;
void lgo_coordinate(const int32_t* n_train, const int32_t* n_valid, const int32_t* n_test,
                    int32_t parts, int32_t batch, int32_t epochs, lgo_steps* out);
int32_t lgo_mode_of_step(const lgo_steps* s, int32_t epochs, int32_t global_batch_id);
int32_t lgo_local_batch_id(const lgo_steps* s, int32_t epochs, int32_t global_batch_id);

#ifdef __cplusplus
}
#endif
#endif
