// harness.cu -- drives the UNMODIFIED reference kernels (compiled where they lie under
// /root/reference/src, see Makefile) below its Server class, so its outputs can pin the oracle
// and the new CUDA path on the B200 box.  TEST INFRASTRUCTURE ONLY; output goes to oracle/_ref/.
//
// The reference's Server::Initialize cannot run here (PCM_Monitor::Init exits without MSR
// access, Server.h:88-98; the loader wants dataset files), so this file mirrors
// GPURunner::Initialize / RunPreSc / RunOnce (Server.cu:169-328) with the reference's own public
// pieces: GPUMemoryNodeStorage / GPUMemoryGraphStorage (Build), GPUCache
// (CandidateSelection / CostModel / FillUp), GPUMemoryPool and the extern "C" operators
// batch_generator_kernel / GPU_Random_Sampling / get_feature_kernel / make_update_plan.
// Nothing of the reference is copied: its sources are compiled in place and linked.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define private public   // the harness reads GPUCache's planning results (QF_/QT_/capacities/shards)
#include "GPUCache.cuh"
#undef private
#include "GPUMemoryPool.cuh"
#include "GPU_Graph_Storage.cuh"
#include "GPU_Node_Storage.cuh"
#include "Kernels.cuh"

struct RefCtx {
    BuildInfo info;
    GPUNodeStorage* node;
    GPUGraphStorage* graph;
    GPUCache* cache;
    GPUMemoryPool* pool;
    cudaStream_t stream;
    cudaStream_t stream2;          // the reference runner's second stream (feature extraction, Server.cu:176-207)
    cudaEvent_t ev[8];
    int32_t* ids[2]; int32_t* labels[2]; int32_t* src_off[2]; int32_t* dst_off[2]; int32_t* nc[2]; int32_t* ec[2];
    float* feats[2];
    int64_t num_ids;
    int32_t n_nodes, dim, batch, f1, f2, pipe;
    int32_t train_step;
};

static void* dmalloc(size_t n) { void* p = nullptr; cudaMalloc(&p, n ? n : 1); return p; }

extern "C" RefCtx* ref_create(const int64_t* indptr, const int32_t* indices, int32_t n_nodes, int64_t n_edges,
                              const float* features, int32_t dim, const int32_t* train_ids, const int32_t* train_labels,
                              int32_t n_train, int32_t batch, int32_t f1, int32_t f2, int64_t cache_memory)
{
    cudaSetDevice(0);
    RefCtx* c = new RefCtx();
    BuildInfo& info = c->info;
    info.partition_count = 1;
    info.shard_to_partition = {0};
    info.shard_to_device = {0};
    info.training_set_ids.assign(1, std::vector<int32_t>(train_ids, train_ids + n_train));
    info.training_labels.assign(1, std::vector<int32_t>(train_labels, train_labels + n_train));
    info.training_set_num = {n_train};
    info.validation_set_ids.assign(1, std::vector<int32_t>(1, train_ids[0]));
    info.validation_labels.assign(1, std::vector<int32_t>(1, train_labels[0]));
    info.validation_set_num = {1};
    info.testing_set_ids.assign(1, std::vector<int32_t>(1, train_ids[0]));
    info.testing_labels.assign(1, std::vector<int32_t>(1, train_labels[0]));
    info.testing_set_num = {1};
    info.total_num_nodes = n_nodes;
    info.int_attr_len = 0;
    info.float_attr_len = dim;
    info.host_int_attrs = nullptr;
    // pinned + mapped host copies, as GPUGraphStore::Load_Graph / Load_Feature allocate them
    cudaHostAlloc(&info.host_float_attrs, (size_t)n_nodes * dim * sizeof(float), cudaHostAllocMapped);
    memcpy(info.host_float_attrs, features, (size_t)n_nodes * dim * sizeof(float));
    cudaHostAlloc(&info.csr_node_index, (size_t)(n_nodes + 1) * sizeof(int64_t), cudaHostAllocMapped);
    memcpy(info.csr_node_index, indptr, (size_t)(n_nodes + 1) * sizeof(int64_t));
    cudaHostAlloc(&info.csr_dst_node_ids, (size_t)n_edges * sizeof(int32_t), cudaHostAllocMapped);
    memcpy(info.csr_dst_node_ids, indices, (size_t)n_edges * sizeof(int32_t));
    info.total_edge_num = n_edges;
    info.cache_edge_num = 0;
    info.epoch = 1;
    info.raw_batch_size = batch;

    c->node = NewGPUMemoryNodeStorage();
    c->node->Build(&info);
    c->graph = NewGPUMemoryGraphStorage();
    c->graph->Build(&info);
    c->train_step = (n_train - 1) / batch;                    // CUDA_IPC_Service.cu:88
    c->cache = new GPUCache();
    c->cache->Initialize(cache_memory, 0, dim, c->train_step, 1);
    c->n_nodes = n_nodes; c->dim = dim; c->batch = batch; c->f1 = f1; c->f2 = f2; c->pipe = 0;
    cudaStreamCreate(&c->stream);
    cudaStreamCreate(&c->stream2);
    for (int i = 0; i < 8; i++) cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);

    // GPURunner::Initialize, Server.cu:183-246
    c->num_ids = (int64_t)batch * (1 + f1 + (int64_t)f1 * f2);
    c->cache->InitializeCacheController(0, n_nodes);
    c->pool = new GPUMemoryPool(2);
    c->pool->SetCacheSearchBuffer((int32_t*)d_alloc_space(c->num_ids * sizeof(int32_t)));
    c->pool->SetAccessedMap((uint32_t*)d_alloc_space((int64_t)(n_nodes / 32 + 1) * sizeof(uint32_t)));
    c->pool->SetPositionMap((int32_t*)d_alloc_space((int64_t)n_nodes * sizeof(int32_t)));
    c->pool->SetAggSrcId((int32_t*)d_alloc_space(c->num_ids * sizeof(int32_t)));
    c->pool->SetAggDstId((int32_t*)d_alloc_space(c->num_ids * sizeof(int32_t)));
    c->pool->SetTmpPartIdx((char*)d_alloc_space(c->num_ids));
    c->pool->SetTmpPartOff((int32_t*)d_alloc_space(c->num_ids * sizeof(int32_t)));
    for (int p = 0; p < 2; p++) {                             // CUDA_IPC_Service.cu:140-215 without the IPC export
        c->ids[p] = (int32_t*)dmalloc(c->num_ids * 4);
        c->labels[p] = (int32_t*)dmalloc((size_t)batch * 4);
        c->src_off[p] = (int32_t*)dmalloc(c->num_ids * 4);
        c->dst_off[p] = (int32_t*)dmalloc(c->num_ids * 4);
        c->nc[p] = (int32_t*)dmalloc(64);
        c->ec[p] = (int32_t*)dmalloc(64);
        c->feats[p] = (float*)dmalloc((size_t)c->num_ids * dim * sizeof(float));
        c->pool->SetSampledIds(c->ids[p], p);
        c->pool->SetLabels(c->labels[p], p);
        c->pool->SetAggSrcOf(c->src_off[p], p);
        c->pool->SetAggDstOf(c->dst_off[p], p);
        c->pool->SetNodeCounter(c->nc[p], p);
        c->pool->SetEdgeCounter(c->ec[p], p);
        c->pool->SetFloatFeatures(c->feats[p], p);
    }
    cudaDeviceSynchronize();
    return c;
}

// GPURunner::RunPreSc (Server.cu:284-299): ops 0, 2, 4, 6 for presampling step `iter`; returns the batch
extern "C" void ref_presample_batch(RefCtx* c, int32_t iter, int32_t* ids, int32_t* src_ids, int32_t* dst_ids, int32_t* nc, int32_t* ec)
{
    c->pool->SetCurrentMode(0);
    c->pool->SetIter(iter);
    batch_generator_kernel(c->stream, c->node, c->cache, c->pool, c->batch, iter, 0, 0, 0);
    GPU_Random_Sampling(c->stream, c->graph, c->cache, c->pool, c->f1, 2, true);
    GPU_Random_Sampling(c->stream, c->graph, c->cache, c->pool, c->f2, 4, true);
    cudaStreamSynchronize(c->stream);
    const int p = c->pipe;
    cudaMemcpy(nc, c->nc[p], 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(ec, c->ec[p], 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(ids, c->ids[p], (size_t)nc[9] * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(src_ids, c->pool->GetAggSrcId(), (size_t)ec[4] * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(dst_ids, c->pool->GetAggDstId(), (size_t)ec[4] * 4, cudaMemcpyDeviceToHost);
    make_update_plan(c->stream, c->graph, c->cache, c->pool, 0, 0);
    cudaStreamSynchronize(c->stream);
}

extern "C" void ref_hotness(RefCtx* c, unsigned long long* node_hot, unsigned long long* topo_hot, int32_t* max_ids)
{
    cudaMemcpy(node_hot, c->cache->cache_controller_[0]->GetNodeAccessedMap(), (size_t)c->n_nodes * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(topo_hot, c->cache->cache_controller_[0]->GetEdgeAccessedMap(), (size_t)c->n_nodes * 8, cudaMemcpyDeviceToHost);
    *max_ids = c->cache->MaxIdNum(0);
}

// Server::PreSc tail (Server.cu:106-108) with a synthetic PCIe counter vector {topo_trans, 0}
extern "C" void ref_plan(RefCtx* c, unsigned long long topo_trans, int32_t* qf, int32_t* qt, int32_t* node_cap, int32_t* edge_cap,
                         float* shard_out /* node_cap*dim, may be NULL */, int64_t shard_rows)
{
    std::vector<uint64_t> counters = {topo_trans, 0};
    c->cache->CandidateSelection(0, c->node, c->graph);
    c->cache->CostModel(0, c->node, c->graph, counters, c->train_step);
    c->cache->FillUp(0, c->node, c->graph);
    cudaDeviceSynchronize();
    cudaMemcpy(qf, c->cache->QF_[0], (size_t)c->n_nodes * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(qt, c->cache->QT_[0], (size_t)c->n_nodes * 4, cudaMemcpyDeviceToHost);
    *node_cap = c->cache->node_capacity_[0];
    *edge_cap = c->cache->edge_capacity_[0];
    if (shard_out) {
        int64_t rows = shard_rows < *node_cap ? shard_rows : *node_cap;
        cudaMemcpy(shard_out, c->cache->Float_Feature_Cache(0), (size_t)rows * c->dim * sizeof(float), cudaMemcpyDeviceToHost);
    }
}

// GPURunner::RunOnce (Server.cu:301-328) for training step `iter`, minus the IPC handshake
extern "C" void ref_train_batch(RefCtx* c, int32_t iter, int32_t* ids, int32_t* labels, int32_t* src_ids, int32_t* dst_ids,
                                int32_t* src_off, int32_t* dst_off, int32_t* nc, int32_t* ec, float* feats)
{
    c->pool->SetCurrentMode(0);
    c->pool->SetIter(iter);
    batch_generator_kernel(c->stream, c->node, c->cache, c->pool, c->batch, iter, 0, 0, 0);
    get_feature_kernel(c->stream, c->cache, c->node, c->pool, 0, 1, true);
    GPU_Random_Sampling(c->stream, c->graph, c->cache, c->pool, c->f1, 2, false);
    get_feature_kernel(c->stream, c->cache, c->node, c->pool, 0, 3, true);
    GPU_Random_Sampling(c->stream, c->graph, c->cache, c->pool, c->f2, 4, false);
    get_feature_kernel(c->stream, c->cache, c->node, c->pool, 0, 5, true);
    make_update_plan(c->stream, c->graph, c->cache, c->pool, 0, 0);
    cudaStreamSynchronize(c->stream);
    const int p = c->pipe;
    cudaMemcpy(nc, c->nc[p], 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(ec, c->ec[p], 64, cudaMemcpyDeviceToHost);
    const int total = nc[9], n_e = ec[4];
    cudaMemcpy(ids, c->ids[p], (size_t)total * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(labels, c->labels[p], (size_t)nc[4] * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(src_ids, c->pool->GetAggSrcId(), (size_t)n_e * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(dst_ids, c->pool->GetAggDstId(), (size_t)n_e * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(src_off, c->src_off[p], (size_t)n_e * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(dst_off, c->dst_off[p], (size_t)n_e * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(feats, c->feats[p], (size_t)total * c->dim * sizeof(float), cudaMemcpyDeviceToHost);
    c->pipe ^= 1;
    c->pool->SetCurrentPipe(c->pipe);
}

extern "C" int64_t ref_capacity(RefCtx* c) { return c->num_ids; }

// Throughput of the reference's own steady-state loop (GPURunner::RunOnce minus the IPC handshake) on this
// GPU: n batches starting at `first_iter`, wall-clock around the whole loop (the reference's operators block
// the host several times per batch anyway, Kernels.cu:605-611, GPUCache.cu:394-395).
#include <chrono>
extern "C" void ref_time_batches(RefCtx* c, int32_t first_iter, int32_t n, double* ms_total, int64_t* edges, int64_t* rows)
{
    *edges = 0; *rows = 0;
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) {
        const int iter = first_iter + i;
        c->pool->SetCurrentMode(0);
        c->pool->SetIter(iter % (c->train_step > 0 ? c->train_step : 1));
        // GPURunner::RunOnce (Server.cu:310-323): operator i runs on stream i % 2, an odd operator (feature extraction,
        // cache update) first makes stream 1 wait for the event of operator i - 1; the batch is complete when the last
        // operator's event has fired (the reference busy-polls it before it touches the next batch)
        cudaStream_t s0 = c->stream, s1 = c->stream2;
        batch_generator_kernel(s0, c->node, c->cache, c->pool, c->batch, iter % (c->train_step > 0 ? c->train_step : 1), 0, 0, 0);
        cudaEventRecord(c->ev[0], s0);
        cudaStreamWaitEvent(s1, c->ev[0], 0);
        get_feature_kernel(s1, c->cache, c->node, c->pool, 0, 1, true);
        GPU_Random_Sampling(s0, c->graph, c->cache, c->pool, c->f1, 2, false);
        cudaEventRecord(c->ev[2], s0);
        cudaStreamWaitEvent(s1, c->ev[2], 0);
        get_feature_kernel(s1, c->cache, c->node, c->pool, 0, 3, true);
        GPU_Random_Sampling(s0, c->graph, c->cache, c->pool, c->f2, 4, false);
        cudaEventRecord(c->ev[4], s0);
        cudaStreamWaitEvent(s1, c->ev[4], 0);
        get_feature_kernel(s1, c->cache, c->node, c->pool, 0, 5, true);
        make_update_plan(s0, c->graph, c->cache, c->pool, 0, 0);
        cudaEventRecord(c->ev[6], s0);
        cudaStreamWaitEvent(s1, c->ev[6], 0);
        update_cache(s1, c->cache, c->node, c->pool, 0, 0);
        cudaEventRecord(c->ev[7], s1);
        cudaEventSynchronize(c->ev[7]);
        c->pipe ^= 1;
        c->pool->SetCurrentPipe(c->pipe);
    }
    cudaDeviceSynchronize();
    *ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int p = 0; p < 2; p++) {   // the last two batches' counters stand in for the per-batch work
        int32_t nc[16], ec[16];
        cudaMemcpy(nc, c->nc[p], 64, cudaMemcpyDeviceToHost);
        cudaMemcpy(ec, c->ec[p], 64, cudaMemcpyDeviceToHost);
        *edges += ec[4]; *rows += nc[9];
    }
    *edges = *edges * n / 2; *rows = *rows * n / 2;
}
