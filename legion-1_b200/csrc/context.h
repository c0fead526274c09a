// context.h -- host-side per-GPU context (the reference's GPURunner + GPUMemoryPool,
// Server.cu:167-364, GPUMemoryPool.cuh:7-208) and the kernel launch prototypes.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace lgn {

// read-only view of the three topology tiers handed to the sampler
struct TopoView {
    const int64_t* base_indptr;                 // full CSR (mapped host or device)
    const int32_t* base_indices;
    const int32_t* slot_of;                     // int32[N]: part*cap+row, -1 = miss, NULL = no cache
    const int64_t* indptr_tab[LGN_MAX_PARTS];   // shard CSRs (local / peer)
    const int32_t* indices_tab[LGN_MAX_PARTS];
    long long cap;
};

// read-only view of the three feature tiers handed to the gather
struct FeatView {
    const float* base;                          // float32[N, D] (mapped host or device)
    const int32_t* slot_of;                     // int32[N]: part*cap+row, -1 = miss, NULL = no cache
    const float* shard_tab[LGN_MAX_PARTS];
    long long cap;
    int32_t my_part;
    int32_t n_parts;
    // compact placement (lgn_place_compact): one 32-byte record per 96 nodes, L2-resident for any N; rows of a class sit in
    // node-id order, so a row index is a prefix count + a popcount (gather.cu: resolve_row)
    const uint4* cmap;                          // NULL: slot_of (or nothing) is bound
    long long n_repl;                           // rows [0, n_repl) of every shard hold the replicated class
    int32_t kg;                                 // GPUs the partitioned class is dealt over
    int32_t identity;                           // 1: every node is resident in shard_tab[my_part] at row = node id (no lookup at all)
};

// One pipeline slot = one independent lane: its own output buffers (the 7 IPC buffers of
// CUDA_IPC_Service.cu:140-215), its own dedup map and scratch, its own gather stream.  Two
// lanes let the latency-bound sampling chain of batch i+1 run while batch i is still in
// flight (the reference shares one bitmap / position map and serialises batches).
struct Pipe {
    int32_t* ids;
    float* features;
    int32_t* labels;
    int32_t* agg_src_off;
    int32_t* agg_dst_off;
    int32_t* nc;
    int32_t* ec;
    // lane-private scratch
    Dedup dedup;               // direct int32[N] map or batch-sized hash table (common.cuh)
    unsigned long long* dedup_tab;   // hash allocation (2^dedup_bits_max entries)
    int32_t* slot_map;         // direct allocation, int32[N]
    int32_t* seed_h;           // hash layout: int32[B], dedup handle of every seed (-1: table full)
    int32_t* agg_src_ids;      // raw ids, int32[capacity]
    int32_t* agg_dst_ids;
    // per-hop scratch, tile t of a hop owns [t * 32 * f, (t + 1) * 32 * f): the hop's valid draws, compacted per tile
    int32_t* draw_h;           // int32[max slots of a hop]: dedup handle of the draw (direct layout: the node id)
    uint16_t* draw_s;          // slot of the draw inside its tile
    int32_t* draw_v;           // payload probed by k_mark
    int32_t* draw_key;         // hash layout: the node id read with the payload
    int32_t* tile_n;           // int32[max tiles of a hop]: valid draws (= edges) of the tile
    int32_t* tile_new;         // new unique nodes of the tile
    int32_t* super_e;          // sums of the two counts per 64 tiles
    int32_t* super_n;
    unsigned long long batch_seq;   // batches started in this slot: generation = 62 - seq % 63
    BatchState* state;         // device
    int32_t* seed_stage;       // device staging for lgn_batch_from_host (ids | labels)
    cudaStream_t gather_stream;
    cudaEvent_t ev_hop[LGN_MAX_HOPS + 2];
    cudaEvent_t ev_end;        // batch_end enqueued on the sampling stream
    cudaEvent_t ev_done;       // everything of the batch in this slot is complete
    bool pending;
    uint32_t external;         // bit i set: wire buffer i (ids, features, labels, agg_src, agg_dst, nc, ec) is caller-owned (lgn_attach_buffers)
    cudaGraphExec_t graph_exec[2][2];   // [with_features][is_presc]: the captured RunOnce / RunPreSc DAG of this slot
    int graph_calls[2][2];
    cudaEvent_t ev_join;                // joins the gather branch back into the captured stream
    cudaStream_t window_stream;   // sampling stream that already carries this lane's L2 access-policy window
};

}  // namespace lgn

struct lgn_ctx {
    lgn_config cfg;
    long long capacity;        // B*(1+f1+f1*f2+...)
    long long max_rows;
    long long max_slots;       // largest F_max*f over hops
    lgn::Pipe pipe[LGN_MAX_LANES];
    int n_lanes;
    int cur_pipe;
    long long max_tiles;       // tiles of the widest hop
    uint32_t* node_hotness;    // u32[N] or NULL
    uint32_t* topo_hotness;
    // bound storage
    const int32_t* seed_ids[3];
    const int32_t* seed_labels[3];
    int32_t seed_count[3];
    lgn::TopoView topo;
    lgn::FeatView feat;
    int use_graphs;            // 1: lgn_run_batch replays a captured CUDA graph per slot (one launch instead of ~20 API calls)
    int l2_persist;            // 1: dedup structures get a persisting L2 access-policy window on the sampling stream
    int dedup_hash;            // 1: hash-table dedup, 0: direct map
    uint32_t dedup_bits_max;   // log2 of the allocated hash table
    int gather_mode;           // 0 = 128-bit LDG/STG warp-per-row, 1 = cp.async.bulk (TMA) thread-per-row, -1 = auto
    int gather_ctas_per_sm;
    int gather_ldg_ctas;       // CTAs per SM of the LDG gather
    int gather_unroll;         // rows in flight per warp of the LDG gather (4, or 2: fewer registers)
    int gather_threads;        // rows in flight per CTA of the bulk-copy gather (<= 256)
    int shared_gather_stream;  // 1: all slots' gathers run back to back on one stream (one saturates HBM already)
    int sample_ctas_per_sm, resolve_ctas_per_sm, end_ctas_per_sm;   // grid caps (CTAs per SM) of the persistent kernels
    int n_sm;
    uint32_t rng_epoch;        // philox counter word 1 of the batches generated from now on (lgn_set_epoch)
    uint32_t rng_step_offset;  // added to the batch counter to form philox counter word 3 (separates train/valid/test streams)
    // optional operator timing (lgn_profile_enable)
    cudaEvent_t* prof_ev;      // 2 events per record
    signed char* prof_kind;
    signed char* prof_pipe;
    int prof_cap, prof_n;
};

namespace lgn {
// sampler.cu
void launch_batch_begin(lgn_ctx* c, cudaStream_t s, const int32_t* ids, const int32_t* labels, int32_t src_off,
                        int32_t count, uint32_t step);
void launch_sample_hop(lgn_ctx* c, cudaStream_t s, int hop, bool presc);
void launch_batch_end(lgn_ctx* c, cudaStream_t s, bool presc);
int sample_items_per_tile();
void sampler_set_carveout(int pct);
// context.cu
void reset_dedup(lgn_ctx* c, Pipe& p, cudaStream_t s);
// gather.cu
void launch_gather(lgn_ctx* c, cudaStream_t s, int segment, int n_segs);
void gather_init_device(const lgn_ctx* c);
const char* gather_kernel_name(const lgn_ctx* c);
void launch_debug_shard_read(lgn_ctx* c, cudaStream_t s, int pipe, long long n_rows, long long rows_per_shard, bool peers_only, uint32_t salt);
void launch_row_copy(const int32_t* order, long long n, long long cap, int kg, int j, long long n_repl, const float* src, int dim,
                     float* dst, int n_sm, cudaStream_t s);
}  // namespace lgn
