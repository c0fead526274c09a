/*
 * legion_b200.h -- C-ABI of the B200-native Legion mini-batch hot path.
 *
 * Plain C: opaque handles, POD structs, raw pointers + sizes, int status codes
 * (0 = LGN_OK, negative = error; lgn_error_string() explains).  No torch / C++
 * types cross this boundary.  Every entry point names the reference interface
 * it replaces (paths relative to the reference checkout, liayan/Legion-1).
 *
 * Pointer conventions: "dev" pointers must be dereferenceable by kernels running
 * on the context's GPU -- local device memory, peer device memory (P2P enabled or
 * opened from an IPC handle) or mapped pinned host memory (UVA zero-copy).
 * All hot-path calls are asynchronous on the caller's stream (a cudaStream_t
 * passed as void*; NULL = legacy default stream) and never synchronise the host.
 */
#ifndef LEGION_B200_H
#define LEGION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGN_MAX_HOPS 5
#define LGN_MAX_PARTS 8        /* MAX_DEVICE, CUDA_IPC_Service.cu:16 */
#define LGN_PIPELINE_DEPTH 2   /* PIPELINE_DEPTH, CUDA_IPC_Service.cu:17: slots visible to the trainer */
#define LGN_MAX_LANES 8        /* independent batch slots a context may own (>= LGN_PIPELINE_DEPTH) */

enum {
    LGN_OK = 0,
    LGN_E_ARG = -1,        /* bad argument */
    LGN_E_CUDA = -2,       /* CUDA runtime error (see lgn_last_cuda_error) */
    LGN_E_STATE = -3,      /* call order violated (e.g. gather before features bound) */
    LGN_E_CAPACITY = -4,   /* batch buffers too small (reference: unchecked, Server.cu:275) */
    LGN_E_SYS = -5         /* shm / semaphore failure */
};

enum { LGN_RNG_MINSTD = 0, LGN_RNG_PHILOX = 1 };
enum { LGN_MODE_TRAIN = 0, LGN_MODE_VALID = 1, LGN_MODE_TEST = 2 };   /* Kernels.cu:10-12 */

const char* lgn_error_string(int code);
const char* lgn_last_cuda_error(void);
int lgn_version(void);

/* ------------------------------------------------------------------ memory
 * replaces d_alloc_space / d_free_space / host_alloc_space / d_copy_2_h /
 * SetGPUDevice / GetGPUDevice (Kernels.cuh:24-46, Kernels.cu:14-64). */
int lgn_set_device(int32_t device);
int lgn_get_device(int32_t* device);
int lgn_device_count(int32_t* count);
int lgn_device_alloc(void** dev_ptr, int64_t bytes);
int lgn_device_free(void* dev_ptr);
/* pinned + mapped host memory; *dev_alias is the UVA pointer kernels may read. */
int lgn_host_alloc_mapped(void** host_ptr, void** dev_alias, int64_t bytes);
int lgn_host_free(void* host_ptr);
int lgn_copy_h2d(void* dev_dst, const void* host_src, int64_t bytes);
int lgn_copy_d2h(void* host_dst, const void* dev_src, int64_t bytes);
int lgn_memset_d(void* dev_dst, int value, int64_t bytes);
int lgn_copy_d2d(void* dev_dst, const void* dev_src, int64_t bytes);   /* any two devices (UVA) */
int lgn_device_synchronize(void);
int lgn_copy_async(void* dst, const void* src, int64_t bytes, void* stream);        /* any direction (UVA), on `stream` */
/* streams for callers that do not link the CUDA runtime themselves (non-blocking; high_priority for sampling lanes) */
int lgn_stream_create(void** stream, int32_t high_priority);
int lgn_stream_destroy(void* stream);
int lgn_stream_synchronize(void* stream);
/* all-pairs cudaDeviceEnablePeerAccess among the first n devices
 * (GPUGraphStore::EnableP2PAccess, GPUGraphStore.cu:145-168). */
int lgn_enable_peer_access(int32_t n_devices);
/* 64-byte CUDA IPC handles for cross-process peer shards / trainer buffers
 * (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, CUDA_IPC_Service.cu:169-175,
 * ipc_cuda_kernel.cu:63-69). */
int lgn_ipc_export(void* dev_ptr, uint8_t handle[64]);
int lgn_ipc_import(const uint8_t handle[64], void** dev_ptr);
int lgn_ipc_close(void* dev_ptr);
/* Cross-process shards through the CUDA virtual-memory-management API (B200 extension; no reference counterpart:
 * the reference is one process driving all GPUs).  lgn_shared_alloc creates device memory on the current device
 * whose shareable handle is a POSIX file descriptor (*fd_out, to be passed to the peer processes over a Unix
 * socket and closed by the caller afterwards) and maps it read/write for the current device; *mapped_bytes is
 * the size rounded up to the allocation granularity, which the importer must pass back.  lgn_shared_import maps
 * such a descriptor into the calling process and grants the CALLING process's current device read/write access
 * (the NVLink peer mapping).  lgn_shared_free unmaps and releases either kind. */
int lgn_shared_alloc(void** dev_ptr, int64_t bytes, int32_t* fd_out, int64_t* mapped_bytes);
int lgn_shared_import(int32_t fd, int64_t mapped_bytes, void** dev_ptr);
int lgn_shared_free(void* dev_ptr);

/* ------------------------------------------------------------------ context
 * One context per GPU = the reference's GPURunner + its GPUMemoryPool
 * (Server.cu:167-364, GPUMemoryPool.cuh:7-208). */
typedef struct lgn_ctx lgn_ctx;

typedef struct {
    int32_t device;                 /* CUDA device ordinal */
    int32_t part;                   /* this GPU's index inside the NVLink clique (0..n_parts-1) */
    int64_t n_nodes;                /* N */
    int32_t feat_dim;               /* D (floats per row) */
    int32_t batch_size;             /* raw batch size B (meta_config field 2) */
    int32_t n_hops;                 /* reference: 2 (Server.cu:68-69) */
    int32_t fanout[LGN_MAX_HOPS];   /* reference: {25,10} */
    int32_t rng_mode;               /* LGN_RNG_* */
    uint64_t rng_seed;              /* philox key */
    int64_t max_feature_rows;       /* 0 = worst case B*(1+f1+f1*f2+..); reference: 1.2*max presampled ids */
    int32_t enable_hotness;         /* allocate the two u32[N] presampling histograms */
    int32_t n_lanes;                /* batch slots ("pipes"): 0 = LGN_PIPELINE_DEPTH (reference); more slots let
                                       more mini-batches be in flight when no trainer throttles the server */
} lgn_config;

int lgn_create(const lgn_config* cfg, lgn_ctx** out);
int lgn_destroy(lgn_ctx* ctx);
/* The per-batch dedup table is allocated for the worst case B*(1+f1+f1*f2+..).  After presampling the real
 * number of unique ids per batch is known (lgn_max_ids); shrinking the table to ~2.5x that keeps it in L2.
 * Overflow is reported through lgn_status (LGN_E_CAPACITY), never silent.  No-op for the direct-map layout. */
int lgn_set_dedup_capacity(lgn_ctx* ctx, int64_t expected_unique);
/* Position of the following batches in the Philox stream (LGN_RNG_PHILOX only): counter = (slot, epoch, hop,
 * step_offset + counter-or-step argument of lgn_batch_generate / lgn_batch_from_host), key = rng_seed.  A server
 * passes the epoch of the global batch id and a per-mode offset (0 / train_step / train_step + valid_step) so that
 * no two mini-batches of a run share a stream; the reference's minstd stream ignores both (it redraws the same
 * neighbourhoods every epoch, Kernels.cu:402-405).  Defaults: 0, 0. */
int lgn_set_epoch(lgn_ctx* ctx, uint32_t epoch, uint32_t step_offset);
/* index of this GPU inside its NVLink clique, once the clique layout is known (PreSc's cache_agg_mode) */
int lgn_set_part(lgn_ctx* ctx, int32_t part);
/* B*(1+f1+f1*f2+...) : per-pipe id / edge buffer capacity (Server.cu:184-196). */
int64_t lgn_capacity(const lgn_ctx* ctx);

/* ------------------------------------------------------------------ storage
 * replaces GPUNodeStorage / GPUGraphStorage pointer tables
 * (GPU_Node_Storage.cuh:24-58, GPU_Graph_Storage.cuh:20-37). */
/* per-mode seed lists of this GPU's partition (GPU_Memory_Node_Storage.cu:52-94) */
int lgn_bind_seeds(lgn_ctx* ctx, int32_t mode, const int32_t* ids_dev, const int32_t* labels_dev, int32_t count);
/* base CSR = slot [partition_count] of the reference tables: the full graph,
 * normally the mapped-host copy (GPU_Memory_Graph_Storage.cu:86-93). */
int lgn_bind_topology(lgn_ctx* ctx, const int64_t* indptr_dev, const int32_t* indices_dev);
/* topology cache shards: slot_of[id] = part*cap + row or -1 (replaces the two
 * cuckoo maps of FindTopo, GPUCache.cu:434-443); NULL slot_of disables it. */
int lgn_bind_topology_cache(lgn_ctx* ctx, int32_t n_parts, const int64_t* const* indptr_tab,
                            const int32_t* const* indices_tab, const int32_t* slot_of_dev, int64_t cap);
/* base feature matrix float32[N,D] (mapped host, GPU_Memory_Node_Storage.cu:22-24) */
int lgn_bind_features(lgn_ctx* ctx, const float* features_dev);
/* feature cache shards: slot_of[id] = part*cap + row or -1 (replaces FindFeat's
 * cuckoo map, GPUCache.cu:387-432); shard_tab = Global_Float_Feature_Cache. */
int lgn_bind_feature_cache(lgn_ctx* ctx, int32_t n_parts, const float* const* shard_tab,
                           const int32_t* slot_of_dev, int64_t cap);
/* the same with a compact placement map (lgn_place_compact).  cmap_dev = NULL with n_repl = n_nodes binds
 * shard_tab[part of this GPU] as the WHOLE feature matrix resident in node-id order: rows are addressed directly,
 * no lookup. */
int lgn_bind_feature_cache_compact(lgn_ctx* ctx, int32_t n_parts, const float* const* shard_tab, const void* cmap_dev,
                                   int64_t n_repl, int64_t cap);

/* ------------------------------------------------------------------ hot path
 * One call per reference operator (Operator.cu:10-123); all asynchronous. */
/* op 0: batch_generator_kernel (Kernels.cu:163-232). Selects the pipe slot. */
int lgn_batch_generate(lgn_ctx* ctx, void* stream, int32_t pipe, int32_t mode, int32_t batch_size, int32_t counter);
/* same, seeds supplied by the caller from (pinned) host memory: H2D inside. */
int lgn_batch_from_host(lgn_ctx* ctx, void* stream, int32_t pipe, const int32_t* seeds_host,
                        const int32_t* labels_host, int32_t count, uint32_t step);
/* ops 2,4,..: GPU_Random_Sampling (Kernels.cu:567-659) for hop = 0..n_hops-1.
 * is_presc: read the base CSR only and count topology hotness (Kernels.cu:468-564). */
int lgn_sample_hop(lgn_ctx* ctx, void* stream, int32_t hop, int32_t is_presc);
/* ops 1,3,5,..: get_feature_kernel (Kernels.cu:707-748) for segment = 0..n_hops. */
int lgn_gather_segment(lgn_ctx* ctx, void* stream, int32_t segment);
/* n_segments adjacent segments in one launch (lgn_run_batch fuses the seeds with hop 1's new nodes) */
int lgn_gather_segments(lgn_ctx* ctx, void* stream, int32_t first_segment, int32_t n_segments);
/* every feature-extraction launch of the current slot's batch, exactly as lgn_run_batch issues them */
int lgn_gather_batch(lgn_ctx* ctx, void* stream);
/* kernels one lgn_batch_generate + lgn_run_batch pair launches (a caller's launch accounting) and the gather variant
 * the bound tiers select */
int32_t lgn_launches_per_batch(const lgn_ctx* ctx, int32_t with_features);
const char* lgn_gather_kernel_name(const lgn_ctx* ctx);
/* op 6/7: make_update_plan / update_cache (Kernels.cu:759-805): node hotness when
 * is_presc (HotnessMeasure, GPUCache.cu:227-235) and scratch reset (ClearPosMap). */
int lgn_finish_batch(lgn_ctx* ctx, void* stream, int32_t is_presc);
/* the whole GPURunner::RunOnce / RunPreSc DAG (Server.cu:284-328) of the selected slot: sampling on `stream`, feature
 * extraction on the slot's own low-priority stream, event-chained per hop.  Replayed as a CUDA graph from the second
 * call on.  Issued eagerly, `stream` is NOT joined with the gather stream at the end; the captured form has to
 * rejoin it (a capture must end on its origin stream), so drive every slot from its own stream: the sampling
 * chains of the other slots then run under this batch's gathers.  A slot's batch is complete when its event fires:
 * lgn_wait_pipe makes `stream` wait for it (device side), lgn_sync_pipe blocks the host
 * (the reference's busy-poll on the last operator event, Server.cu:318-323). */
int lgn_run_batch(lgn_ctx* ctx, void* stream, int32_t with_features, int32_t is_presc);
/* operator-by-operator calls act on the slot chosen by the last batch_generate / select */
int lgn_select_pipe(lgn_ctx* ctx, int32_t pipe);
int lgn_wait_pipe(lgn_ctx* ctx, void* stream, int32_t pipe);
int lgn_sync_pipe(lgn_ctx* ctx, int32_t pipe);
/* lgn_sync_pipe + the slot's sticky device status (0 or LGN_E_CAPACITY) in one call */
int lgn_sync_pipe_status(lgn_ctx* ctx, int32_t pipe);

/* ------------------------------------------------------------------ results */
typedef struct {
    int32_t* ids;           /* int32[capacity]   sampled_ids  (IPC handle 0) */
    float* features;        /* f32[max_rows, D]               (IPC handle 1) */
    int32_t* labels;        /* int32[B]                       (IPC handle 2) */
    int32_t* agg_src;       /* int32[capacity]   local index of sampled neighbour (IPC handle 3) */
    int32_t* agg_dst;       /* int32[capacity]   local index of frontier node     (IPC handle 4) */
    int32_t* node_counter;  /* int32[16]                      (IPC handle 5) */
    int32_t* edge_counter;  /* int32[16]                      (IPC handle 6) */
    int32_t* agg_src_ids;   /* raw ids (GPUMemoryPool::GetAggSrcId); one array per slot here (the reference shares one) */
    int32_t* agg_dst_ids;   /* raw ids (GPUMemoryPool::GetAggDstId) */
    int64_t capacity;
    int64_t max_rows;
} lgn_batch_view;
int lgn_batch_buffers(lgn_ctx* ctx, int32_t pipe, lgn_batch_view* out);
/* the reverse: the caller owns the wire buffers of a slot (GPUMemoryPool::Set*, GPUMemoryPool.cuh:92-160; the reference
 * runner allocates them itself or takes them from IPCEnv, Server.cu:217-283).  Non-NULL members of *v replace the
 * slot's buffers; capacity / max_rows state their sizes; agg_src_ids / agg_dst_ids are ignored (lane-private). */
int lgn_attach_buffers(lgn_ctx* ctx, int32_t pipe, const lgn_batch_view* v);
/* D2H of both counter blocks on `stream` + stream sync: what ipc_service.get_next
 * does on the trainer side (ipc_cuda_kernel.cu:192-193). */
int lgn_read_counters(lgn_ctx* ctx, void* stream, int32_t pipe, int32_t nc[16], int32_t ec[16]);
/* rows served per tier by the gathers since the last reset: [local, peer, host] */
int lgn_tier_counts(lgn_ctx* ctx, void* stream, int64_t out[3], int32_t reset);
/* sticky device-side status: 0 or LGN_E_CAPACITY */
int lgn_status(lgn_ctx* ctx, void* stream);

/* CUDA-event timing of the operators (bench.py's live roofline): every operator call records an
 * event pair on its own stream while enabled.  kinds: 0 batch begin, 1 sampling hop (3 kernels),
 * 2 feature gather, 3 batch end.  collect() synchronises, sums and resets. */
int lgn_profile_enable(lgn_ctx* ctx, int32_t max_records);
int lgn_profile_collect(lgn_ctx* ctx, double ms_by_kind[4], int64_t calls_by_kind[4]);
/* raw timeline: up to max_records rows of {kind, slot, begin_ms, end_ms} relative to the first record; returns
 * the number of rows through *n_out.  Does not reset. */
int lgn_profile_timeline(lgn_ctx* ctx, double* rows4, int32_t max_records, int32_t* n_out);
/* diagnostic: gathers n_rows pseudo-random rows (row index < rows_per_shard, uniformly spread over the bound
 * cache shards; peers_only != 0 skips this GPU's own shard) into `slot`'s feature buffer with the plain 128-bit
 * LDG loop of tools/peer_probe.cu -- no id list, no slot table, no cache hints -- `repeats` times, and returns
 * the average launch time.  Separates "the mapping is slow" from "the gather kernel / its data is slow" when
 * the NVLink tier under-performs.  No reference counterpart. */
int lgn_debug_shard_read(lgn_ctx* ctx, void* stream, int32_t slot, int64_t n_rows, int64_t rows_per_shard, int32_t peers_only,
                         int32_t repeats, double* avg_ms);

/* ------------------------------------------------------------------ planner
 * presampling statistics and cache construction (GPUCache.cu:578-826). */
int lgn_hotness(lgn_ctx* ctx, uint32_t** node_hotness_dev, uint32_t** topo_hotness_dev);
int32_t lgn_max_ids(lgn_ctx* ctx, void* stream);
/* frontier items expanded and edges sampled since the last reset: the analytic stand-in for the PCIe
 * read-transaction counters the reference takes from Intel PCM (Server.cu:84-108, GPUCache.cu:675):
 * one 64-byte transaction per indptr pair and per neighbour read over UVA. */
int lgn_sampling_totals(lgn_ctx* ctx, void* stream, int64_t out[2], int32_t reset);
/* dst[i] += src[i]; src may live on a peer GPU (aggregate_access, GPUCache.cu:44-48, 624-627) */
int lgn_accumulate_u32(uint32_t* dst_dev, const uint32_t* src_dev, int64_t n, void* stream);  /* PreSCCacheController::MaxIdNum (GPUCache.cu:294-296) */
/* order[i] = id of rank i under (count desc, id asc): CandidateSelection's
 * sort_by_key (GPUCache.cu:630-631). sorted_counts may be NULL. */
int lgn_hot_order(const uint32_t* counts_dev, int64_t n, int32_t* order_dev, uint32_t* sorted_counts_dev, void* stream);
/* slot_of[id] = (i%kg)*cap + i/kg for rank i < min(cap*kg, n) else -1 (InitPair, GPUCache.cu:103-108) */
int lgn_place(const int32_t* order_dev, int64_t n, int64_t cap, int32_t kg, int32_t* slot_of_dev, void* stream);
/* shard j row r <- features[order[r*kg+j]] (FeatFillUp, GPUCache.cu:200-205) */
int lgn_fill_feature_shard(const int32_t* order_dev, int64_t n, int64_t cap, int32_t kg, int32_t j,
                           const float* features_dev, int32_t dim, float* shard_dev, void* stream);
/* B200 extension (no reference counterpart): hybrid placement.  The n_repl hottest ranks are
 * REPLICATED on every GPU of the clique (rows [0, n_repl) of each shard, served from local HBM), the next
 * (cap - n_repl) * kg ranks are partitioned as above into rows [n_repl, cap).  n_repl = 0 is the reference
 * placement; n_repl = cap replicates everything (the reference's cache_agg_mode 0).  The slot map differs per
 * GPU: replicated ranks point at `my_part`. */
int lgn_place_hybrid(const int32_t* order_dev, int64_t n, int64_t cap, int32_t kg, int64_t n_repl, int32_t my_part,
                     int32_t* slot_of_dev, void* stream);
int lgn_fill_feature_shard_hybrid(const int32_t* order_dev, int64_t n, int64_t cap, int32_t kg, int32_t j, int64_t n_repl,
                                  const float* features_dev, int32_t dim, float* shard_dev, void* stream);
/* B200 extension: COMPACT placement map.  Same three classes as the hybrid placement (the n_repl hottest ranks of
 * `order` replicated on every GPU, the next n_part ranks partitioned over kg GPUs, the rest on the host tier), but the
 * rows of a class are stored in NODE-ID order, so the map needs no row number: one 32-byte record per LGN_CMAP_NODES
 * nodes { u32 replicated nodes before the record, u32 partitioned nodes before it, 3 x u32 "replicated" bits,
 * 3 x u32 "partitioned" bits }.  A replicated node's row is its rank among replicated nodes (prefix + popcount);
 * a partitioned node with rank q among partitioned nodes lives on GPU q % kg, row n_repl + q / kg.  The map of
 * papers100M is 37 MB and stays in L2, where int32 slot_of[N] (444 MB) costs a random DRAM access per gathered row.
 * Replaces FindFeat's cuckoo probe (GPUCache.cu:387-432) like lgn_place; which rows are cached where is still
 * decided by hotness rank (GPUCache.cu:103-108), only the order inside a shard differs.  The map is the same on
 * every GPU of the clique. */
#define LGN_CMAP_NODES 96
int64_t lgn_cmap_bytes(int64_t n_nodes);
int lgn_place_compact(const int32_t* order_dev, int64_t n, int64_t n_repl, int64_t n_part, void* cmap_dev, void* stream);
/* shard j (height cap >= n_repl + ceil(n_part / kg)) of that placement: a streaming pass over the feature matrix */
int lgn_fill_feature_shard_compact(const void* cmap_dev, int64_t n, int64_t n_repl, int32_t kg, int32_t j,
                                   const float* features_dev, int32_t dim, float* shard_dev, int64_t cap, void* stream);
/* topology shard j (GraphCache, GPU_Memory_Graph_Storage.cu:98-133). Pass
 * indices_out_dev = NULL to size it: *n_indices receives the count (host sync). */
int lgn_fill_topo_shard(const int32_t* order_dev, int64_t n, int64_t cap, int32_t kg, int32_t j,
                        const int64_t* indptr_dev, const int32_t* indices_dev,
                        int64_t* indptr_out_dev, int32_t* indices_out_dev, int64_t* n_indices, void* stream);
/* CostModel (GPUCache.cu:661-767); sorted counts / order are device arrays. */
int lgn_cost_model(const uint32_t* af_sorted_dev, const uint32_t* at_sorted_dev, const int32_t* qt_dev,
                   const int64_t* indptr_dev, int64_t n, int32_t dim, int64_t cache_memory, int32_t kg,
                   uint64_t topo_trans, const int32_t* max_ids, int32_t train_step,
                   int32_t* node_capacity, int32_t* edge_capacity, void* stream);
/* B200 placement model (no reference counterpart): how many of the hottest feature rows to replicate on
 * every GPU of the clique before the rest is partitioned and what stays on the host, given the per-GPU byte budget for
 * features and the three tier bandwidths (any common unit).  af_sorted_dev = presampled hotness in hot order (the
 * sorted counts of lgn_hot_order); `prior` = pseudo-count added to every row (rows one presampling epoch never saw
 * are still read now and then: with prior > 0 they are kept off the host tier while the clique has room).  Outputs feed lgn_place_hybrid / lgn_fill_feature_shard_hybrid: *n_repl rows are
 * replicated, *cap is the shard height. */
int lgn_plan_hybrid(const uint32_t* af_sorted_dev, int64_t n, int32_t dim, int64_t budget_bytes, int32_t kg,
                    double bw_local, double bw_peer, double bw_host, double prior, int64_t* n_repl, int64_t* cap,
                    double* est_cost, void* stream);

/* ------------------------------------------------------------------ collective
 * The path's one collective: the sum of the per-GPU hotness histograms before planning.  Replaces the leader's P2P
 * reads of aggregate_access (GPUCache.cu:44-48, 624-647) with an NCCL all-reduce of u32[N]; NCCL is
 * resolved with dlopen at first use (LGN_E_SYS when it is absent). */
int lgn_comm_available(void);
/* one process drives the GPUs of the clique (the `legion` server): bufs[i] is device memory of devices[i]; blocks
 * until every buffer holds the element-wise sum */
int lgn_allreduce_u32_devices(int32_t n_devices, const int32_t* devices, uint32_t* const* bufs, int64_t count);
/* one process per GPU: rank 0 makes the 128-byte NCCL unique id, the caller ships it to the other ranks, every rank
 * creates its communicator with its GPU current, then all-reduces in place on a stream */
typedef struct lgn_comm lgn_comm;
int lgn_comm_unique_id(uint8_t id[128]);
int lgn_comm_create(int32_t rank, int32_t world, const uint8_t id[128], lgn_comm** out);
int lgn_comm_allreduce_u32(lgn_comm* comm, uint32_t* buf_dev, int64_t count, void* stream);
int lgn_comm_destroy(lgn_comm* comm);

/* ------------------------------------------------------------------ IPC wire format
 * server side of CUDA_IPC_Service (CUDA_IPC_Service.cu:34-359): POSIX shm
 * "simpleIPCshm" {int32 steps[3]; handle[8][2][7]}, sems sem_r_/sem_w_<dev>_<pipe>. */
typedef struct lgn_ipc_server lgn_ipc_server;
int lgn_ipc_server_create(int32_t n_devices, const int32_t steps[3], lgn_ipc_server** out);
int lgn_ipc_server_publish(lgn_ipc_server* s, int32_t device, lgn_ctx* ctx, int32_t with_features);
int lgn_ipc_server_wait(lgn_ipc_server* s, int32_t device, int32_t pipe);   /* IPCWait  (sem_r) */
int lgn_ipc_server_post(lgn_ipc_server* s, int32_t device, int32_t pipe);   /* IPCPost  (sem_w) */
int lgn_ipc_server_destroy(lgn_ipc_server* s);
/* trainer side (pytorch_extension/ipc_cuda_kernel.cu:38-176) */
typedef struct lgn_ipc_client lgn_ipc_client;
int lgn_ipc_client_open(int32_t device, lgn_ipc_client** out);
int lgn_ipc_client_steps(lgn_ipc_client* c, int32_t steps[3]);
/* Wait() + counter D2H; fills the 7 device pointers of the current pipe */
int lgn_ipc_client_next(lgn_ipc_client* c, void* ptrs[7], int32_t nc[16], int32_t ec[16]);
int lgn_ipc_client_release(lgn_ipc_client* c);                              /* Post() + pipe flip */
int lgn_ipc_client_close(lgn_ipc_client* c);

/* step arithmetic of IPCEnv::Coordinate / GetCurrentMode / GetLocalBatchId
 * (CUDA_IPC_Service.cu:66-134, 219-259). */
typedef struct {
    int32_t train_step, valid_step, test_step, max_step;
    int32_t valid_batch[LGN_MAX_PARTS], test_batch[LGN_MAX_PARTS];
} lgn_steps;
int lgn_coordinate(const int32_t* n_train, const int32_t* n_valid, const int32_t* n_test, int32_t parts,
                   int32_t batch, int32_t epochs, lgn_steps* out);
int32_t lgn_mode_of_step(const lgn_steps* s, int32_t epochs, int32_t global_batch_id);
int32_t lgn_local_batch_id(const lgn_steps* s, int32_t epochs, int32_t global_batch_id);

#ifdef __cplusplus
}
#endif
#endif
