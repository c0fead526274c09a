// reference_abi.cpp -- the reference's kernel-library entry points (reference_abi.h) as thin calls into the C-ABI.
// Host code only; every GPU action goes through include/legion_b200.h.  Errors are fatal, as in the reference
// (cudaCheckError -> exit, Kernels.cuh:14-22).
#include "reference_abi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <sstream>
#include <string>

#include "../../../include/legion_b200.h"

#define ABI_DIE(rc, what)                                                                                          \
    do {                                                                                                           \
        int rc_ = (rc);                                                                                            \
        if (rc_ != 0) {                                                                                            \
            fprintf(stderr, "legion_b200 (reference ABI): %s failed: %s %s\n", what, lgn_error_string(rc_), lgn_last_cuda_error()); \
            exit(EXIT_FAILURE);                                                                                    \
        }                                                                                                          \
    } while (0)

namespace {

struct DevState {
    lgn_ctx* ctx = nullptr;
    bool seeds = false, topo = false, feat = false;
    int n_hops = 2;
    int fanout[LGN_MAX_HOPS] = {25, 10, 0, 0, 0};
};
DevState g_dev[LGN_MAX_PARTS];
std::mutex g_mu;
int32_t g_raw_batch = 0;

void read_fanout(DevState& d)
{
    if (const char* e = getenv("LEGION_FANOUT")) {        // the reference hard-codes {25, 10} (Server.cu:68-69)
        std::stringstream ss(e);
        std::string tok;
        int n = 0;
        while (std::getline(ss, tok, ',') && n < LGN_MAX_HOPS) d.fanout[n++] = atoi(tok.c_str());
        if (n > 0) d.n_hops = n;
    }
}

lgn_ctx* ensure_ctx(int dev, int64_t n_nodes, int dim, int batch)
{
    std::lock_guard<std::mutex> g(g_mu);
    if (dev < 0 || dev >= LGN_MAX_PARTS) { fprintf(stderr, "legion_b200 (reference ABI): device %d out of range\n", dev); exit(EXIT_FAILURE); }
    DevState& d = g_dev[dev];
    if (d.ctx) return d.ctx;
    read_fanout(d);
    lgn_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = dev; cfg.part = 0; cfg.n_nodes = n_nodes; cfg.feat_dim = dim;
    cfg.batch_size = batch > g_raw_batch ? batch : g_raw_batch;      // the reference sizes everything from the raw batch (Server.cu:184-196)
    cfg.n_hops = d.n_hops;
    for (int h = 0; h < d.n_hops; h++) cfg.fanout[h] = d.fanout[h];
    const char* rng = getenv("LEGION_RNG");
    const char* seed = getenv("LEGION_SEED");
    cfg.rng_mode = (rng && !strcmp(rng, "philox")) ? LGN_RNG_PHILOX : LGN_RNG_MINSTD;
    cfg.rng_seed = seed ? strtoull(seed, nullptr, 0) : 0;
    cfg.enable_hotness = 1;
    cfg.n_lanes = LGN_PIPELINE_DEPTH;
    ABI_DIE(lgn_create(&cfg, &d.ctx), "lgn_create");
    return d.ctx;
}

// ---- storages (GPU_Memory_Node_Storage.cu:10-120, GPU_Memory_Graph_Storage.cu:37-96) --------------------------
class MemoryNodeStorage : public GPUNodeStorage {
public:
    void Build(BuildInfo* info) override
    {
        n_ = info->total_num_nodes; dim_ = info->float_attr_len; parts_ = info->partition_count;
        g_raw_batch = info->raw_batch_size;
        feat_ = info->host_float_attrs;                 // already the device alias when it came from host_alloc_space
        for (int m = 0; m < 3; m++) { ids_[m].assign(parts_, nullptr); lab_[m].assign(parts_, nullptr); cnt_[m].assign(parts_, 0); }
        const std::vector<std::vector<int32_t>>* src_ids[3] = {&info->training_set_ids, &info->validation_set_ids, &info->testing_set_ids};
        const std::vector<std::vector<int32_t>>* src_lab[3] = {&info->training_labels, &info->validation_labels, &info->testing_labels};
        int cur = 0;
        lgn_get_device(&cur);
        for (size_t i = 0; i < info->shard_to_partition.size(); i++) {
            const int part = info->shard_to_partition[i], dev = info->shard_to_device[i];
            ABI_DIE(lgn_set_device(dev), "SetGPUDevice");
            for (int m = 0; m < 3; m++) {
                const std::vector<int32_t>& a = (*src_ids[m])[part];
                const std::vector<int32_t>& b = (*src_lab[m])[part];
                void *di = nullptr, *dl = nullptr;
                ABI_DIE(lgn_device_alloc(&di, (int64_t)a.size() * 4), "alloc ids");
                ABI_DIE(lgn_device_alloc(&dl, (int64_t)a.size() * 4), "alloc labels");
                if (!a.empty()) { ABI_DIE(lgn_copy_h2d(di, a.data(), (int64_t)a.size() * 4), "h2d"); ABI_DIE(lgn_copy_h2d(dl, b.data(), (int64_t)a.size() * 4), "h2d"); }
                ids_[m][part] = (int32_t*)di; lab_[m][part] = (int32_t*)dl; cnt_[m][part] = (int32_t)a.size();
            }
        }
        lgn_set_device(cur);
    }
    void Finalize() override
    {
        for (int m = 0; m < 3; m++) for (int p = 0; p < parts_; p++) { lgn_device_free(ids_[m][p]); lgn_device_free(lab_[m][p]); ids_[m][p] = lab_[m][p] = nullptr; }
    }
    int32_t* GetTrainingSetIds(int32_t p) const override { return ids_[0][p]; }
    int32_t* GetValidationSetIds(int32_t p) const override { return ids_[1][p]; }
    int32_t* GetTestingSetIds(int32_t p) const override { return ids_[2][p]; }
    int32_t* GetTrainingLabels(int32_t p) const override { return lab_[0][p]; }
    int32_t* GetValidationLabels(int32_t p) const override { return lab_[1][p]; }
    int32_t* GetTestingLabels(int32_t p) const override { return lab_[2][p]; }
    int32_t TrainingSetSize(int32_t p) const override { return cnt_[0][p]; }
    int32_t ValidationSetSize(int32_t p) const override { return cnt_[1][p]; }
    int32_t TestingSetSize(int32_t p) const override { return cnt_[2][p]; }
    int32_t TotalNodeNum() const override { return n_; }
    float* GetAllFloatAttr() const override { return feat_; }
    int32_t GetFloatAttrLen() const override { return dim_; }

private:
    int32_t n_ = 0, dim_ = 0, parts_ = 0;
    float* feat_ = nullptr;
    std::vector<int32_t*> ids_[3], lab_[3];
    std::vector<int32_t> cnt_[3];
};

class MemoryGraphStorage : public GPUGraphStorage {
public:
    void Build(BuildInfo* info) override
    {
        parts_ = info->partition_count; indptr_ = info->csr_node_index; indices_ = info->csr_dst_node_ids; edges_ = info->total_edge_num;
    }
    void GraphCache(int32_t*, int32_t, int32_t, int32_t) override {}      // shards are built by GPUCache::FillUp here
    void Finalize() override {}
    int32_t GetPartitionCount() const override { return parts_; }
    int64_t* GetCSRNodeIndexCPU() const override { return indptr_; }
    int32_t* GetCSRNodeMatrixCPU() const override { return indices_; }
    int64_t Src_Size(int32_t) const override { return 0; }
    int64_t Dst_Size(int32_t) const override { return edges_; }

private:
    int32_t parts_ = 0;
    int64_t* indptr_ = nullptr;
    int32_t* indices_ = nullptr;
    int64_t edges_ = 0;
};

}  // namespace

lgn_ctx* LegionContextOfDevice(int32_t dev) { return dev >= 0 && dev < LGN_MAX_PARTS ? g_dev[dev].ctx : nullptr; }
void LegionReleaseContexts()
{
    for (DevState& d : g_dev) { if (d.ctx) lgn_destroy(d.ctx); d = DevState(); }
}

// ---- GPUMemoryPool -------------------------------------------------------------------------------------------
GPUMemoryPool::GPUMemoryPool(int32_t pipeline_depth) : depth_(pipeline_depth > kDepth ? kDepth : pipeline_depth) {}
int32_t* GPUMemoryPool::GetAggSrcId()
{
    lgn_batch_view v;
    if (!ctx || lgn_batch_buffers(ctx, pipe_, &v)) return nullptr;
    return v.agg_src_ids;
}
int32_t* GPUMemoryPool::GetAggDstId()
{
    lgn_batch_view v;
    if (!ctx || lgn_batch_buffers(ctx, pipe_, &v)) return nullptr;
    return v.agg_dst_ids;
}
void GPUMemoryPool::Sync()
{
    if (!ctx || !dirty_) return;
    for (int p = 0; p < depth_; p++) {
        lgn_batch_view v;
        memset(&v, 0, sizeof(v));
        v.ids = sampled_ids_[p]; v.features = float_features_[p]; v.labels = labels_[p];
        v.agg_src = agg_src_off_[p]; v.agg_dst = agg_dst_off_[p]; v.node_counter = node_counter_[p]; v.edge_counter = edge_counter_[p];
        v.capacity = num_ids_ > 0 ? num_ids_ : lgn_capacity(ctx);
        v.max_rows = feature_rows_ > 0 ? feature_rows_ : lgn_capacity(ctx);
        ABI_DIE(lgn_attach_buffers(ctx, p, &v), "GPUMemoryPool::Set* (lgn_attach_buffers)");
    }
    dirty_ = false;
}

// ---- C entry points (Kernels.cu:14-64, 163-232, 567-659, 707-805) --------------------------------------------
extern "C" {

void* d_alloc_space(int64_t num_bytes) { void* p = nullptr; ABI_DIE(lgn_device_alloc(&p, num_bytes), "d_alloc_space"); return p; }
void* d_alloc_space_managed(unsigned int num_bytes) { return d_alloc_space((int64_t)num_bytes); }   // nothing on this path relies on migration
void d_copy_2_h(void* h_ptr, void* d_ptr, unsigned int num_bytes) { ABI_DIE(lgn_copy_d2h(h_ptr, d_ptr, (int64_t)num_bytes), "d_copy_2_h"); }
void d_free_space(void* d_ptr) { ABI_DIE(lgn_device_free(d_ptr), "d_free_space"); }
void SetGPUDevice(int32_t shard_id) { ABI_DIE(lgn_set_device(shard_id), "SetGPUDevice"); }
int32_t GetGPUDevice() { int32_t d = 0; ABI_DIE(lgn_get_device(&d), "GetGPUDevice"); return d; }
void* host_alloc_space(unsigned int num_bytes)
{
    void *h = nullptr, *d = nullptr;
    ABI_DIE(lgn_host_alloc_mapped(&h, &d, (int64_t)num_bytes), "host_alloc_space");
    return d;
}

void batch_generator_kernel(cudaStream_t strm_hdl, GPUNodeStorage* noder, GPUCache*, GPUMemoryPool* memorypool, int32_t batch_size,
                            int32_t counter, int32_t, int32_t dev_id, int32_t mode)
{
    lgn_ctx* ctx = ensure_ctx(dev_id, noder->TotalNodeNum(), noder->GetFloatAttrLen(), batch_size);
    DevState& d = g_dev[dev_id];
    if (!d.seeds) {      // the reference indexes the seed sets by dev_id (Kernels.cu:178-192)
        ABI_DIE(lgn_bind_seeds(ctx, LGN_MODE_TRAIN, noder->GetTrainingSetIds(dev_id), noder->GetTrainingLabels(dev_id), noder->TrainingSetSize(dev_id)), "bind train seeds");
        ABI_DIE(lgn_bind_seeds(ctx, LGN_MODE_VALID, noder->GetValidationSetIds(dev_id), noder->GetValidationLabels(dev_id), noder->ValidationSetSize(dev_id)), "bind valid seeds");
        ABI_DIE(lgn_bind_seeds(ctx, LGN_MODE_TEST, noder->GetTestingSetIds(dev_id), noder->GetTestingLabels(dev_id), noder->TestingSetSize(dev_id)), "bind test seeds");
        d.seeds = true;
    }
    memorypool->ctx = ctx;
    memorypool->Sync();
    ABI_DIE(lgn_batch_generate(ctx, strm_hdl, memorypool->pipe(), mode, batch_size, counter), "batch_generator_kernel");
}

void GPU_Random_Sampling(cudaStream_t strm_hdl, GPUGraphStorage* graph, GPUCache*, GPUMemoryPool* memorypool, int32_t count, int32_t op_id,
                         bool is_presc)
{
    lgn_ctx* ctx = memorypool->ctx;
    if (!ctx) { fprintf(stderr, "legion_b200 (reference ABI): GPU_Random_Sampling before batch_generator_kernel\n"); exit(EXIT_FAILURE); }
    int dev = 0;
    for (int i = 0; i < LGN_MAX_PARTS; i++) if (g_dev[i].ctx == ctx) dev = i;
    DevState& d = g_dev[dev];
    if (!d.topo) {
        ABI_DIE(lgn_bind_topology(ctx, graph->GetCSRNodeIndexCPU(), graph->GetCSRNodeMatrixCPU()), "bind topology");
        d.topo = true;
    }
    const int hop = op_id / 2 - 1;                       // op ids 2, 4, .. (Server.cu:198-207)
    if (hop < 0 || hop >= d.n_hops || count != d.fanout[hop]) {
        fprintf(stderr, "legion_b200 (reference ABI): sampler op %d with fan-out %d does not match LEGION_FANOUT\n", op_id, count);
        exit(EXIT_FAILURE);
    }
    ABI_DIE(lgn_select_pipe(ctx, memorypool->pipe()), "lgn_select_pipe");
    ABI_DIE(lgn_sample_hop(ctx, strm_hdl, hop, is_presc ? 1 : 0), "GPU_Random_Sampling");
}

void get_feature_kernel(cudaStream_t strm_hdl, GPUCache*, GPUNodeStorage* noder, GPUMemoryPool* memorypool, int32_t dev_id, int32_t op_id, bool)
{
    lgn_ctx* ctx = memorypool->ctx;
    if (!ctx) { fprintf(stderr, "legion_b200 (reference ABI): get_feature_kernel before batch_generator_kernel\n"); exit(EXIT_FAILURE); }
    DevState& d = g_dev[dev_id];
    if (!d.feat) {
        ABI_DIE(lgn_bind_features(ctx, noder->GetAllFloatAttr()), "bind features");
        d.feat = true;
    }
    ABI_DIE(lgn_select_pipe(ctx, memorypool->pipe()), "lgn_select_pipe");
    ABI_DIE(lgn_gather_segment(ctx, strm_hdl, (op_id - 1) / 2), "get_feature_kernel");      // op ids 1, 3, 5
}

void make_update_plan(cudaStream_t strm_hdl, GPUGraphStorage*, GPUCache* cache, GPUMemoryPool* memorypool, int32_t, int32_t mode)
{
    lgn_ctx* ctx = memorypool->ctx;
    if (!ctx) return;
    ABI_DIE(lgn_select_pipe(ctx, memorypool->pipe()), "lgn_select_pipe");
    // the reference profiles accesses of TRAIN batches until the cache is built (CacheProfiling, GPUCache.cu:275-303)
    const int presc = (mode == LGN_MODE_TRAIN && cache && !cache->filled()) ? 1 : 0;
    ABI_DIE(lgn_finish_batch(ctx, strm_hdl, presc), "make_update_plan");
}

void update_cache(cudaStream_t, GPUCache*, GPUNodeStorage*, GPUMemoryPool*, int32_t, int32_t) {}

GPUGraphStorage* NewGPUMemoryGraphStorage() { return new MemoryGraphStorage(); }
GPUNodeStorage* NewGPUMemoryNodeStorage() { return new MemoryNodeStorage(); }

}  // extern "C"

// ---- GPUCache (GPUCache.cu:578-826) ----------------------------------------------------------------------------
void GPUCache::Initialize(int64_t cache_memory, int32_t, int32_t float_attr_len, int32_t train_step, int32_t device_count)
{
    cache_memory_ = cache_memory; dim_ = float_attr_len; train_step_ = train_step; n_dev_ = device_count;
}
void GPUCache::InitializeCacheController(int32_t, int32_t) {}
void GPUCache::Finalize(int32_t dev_id)
{
    if (dev_id != 0) return;
    for (void* p : owned_) lgn_device_free(p);
    owned_.clear();
}
int32_t GPUCache::NodeCapacity(int32_t) { return node_cap_; }
float* GPUCache::Float_Feature_Cache(int32_t dev_id) { return dev_id < (int32_t)shards_.size() ? (float*)shards_[dev_id] : nullptr; }
int32_t GPUCache::MaxIdNum(int32_t dev_id) { lgn_ctx* c = LegionContextOfDevice(dev_id); return c ? lgn_max_ids(c, nullptr) : 0; }

void GPUCache::CandidateSelection(int cache_agg_mode, GPUNodeStorage* noder, GPUGraphStorage*)
{
    kg_ = cache_agg_mode == 1 ? 2 : cache_agg_mode == 2 ? 4 : cache_agg_mode == 3 ? 8 : 1;      // GPUCache.cu:593-607
    if (kg_ > n_dev_) kg_ = n_dev_;
    const int64_t N = noder->TotalNodeNum();
    const int kc = n_dev_ / kg_;
    qf_.assign(kc, nullptr); qt_.assign(kc, nullptr); af_.assign(kc, nullptr); at_.assign(kc, nullptr);
    for (int c = 0; c < kc; c++) {
        const int lead = c * kg_;
        std::vector<int32_t> devs;
        std::vector<uint32_t*> nptr, tptr;
        for (int j = 0; j < kg_; j++) {
            lgn_ctx* ctx = LegionContextOfDevice(lead + j);
            if (!ctx) { fprintf(stderr, "legion_b200 (reference ABI): CandidateSelection before presampling on GPU %d\n", lead + j); exit(EXIT_FAILURE); }
            uint32_t *a = nullptr, *b = nullptr;
            ABI_DIE(lgn_hotness(ctx, &a, &b), "lgn_hotness");
            devs.push_back(lead + j); nptr.push_back(a); tptr.push_back(b);
        }
        ABI_DIE(lgn_set_device(lead), "SetGPUDevice");
        if (kg_ > 1 && lgn_comm_available()) {          // sum over the clique: NCCL all-reduce (aggregate_access, GPUCache.cu:624-647)
            ABI_DIE(lgn_allreduce_u32_devices(kg_, devs.data(), nptr.data(), N), "ncclAllReduce");
            ABI_DIE(lgn_allreduce_u32_devices(kg_, devs.data(), tptr.data(), N), "ncclAllReduce");
        } else {
            for (int j = 1; j < kg_; j++) { ABI_DIE(lgn_accumulate_u32(nptr[0], nptr[j], N, nullptr), "aggregate_access"); ABI_DIE(lgn_accumulate_u32(tptr[0], tptr[j], N, nullptr), "aggregate_access"); }
        }
        ABI_DIE(lgn_device_synchronize(), "sync");
        for (void** p : {&qf_[c], &qt_[c], &af_[c], &at_[c]}) { ABI_DIE(lgn_device_alloc(p, N * 4), "alloc"); owned_.push_back(*p); }
        ABI_DIE(lgn_hot_order(nptr[0], N, (int32_t*)qf_[c], (uint32_t*)af_[c], nullptr), "hot order (features)");
        ABI_DIE(lgn_hot_order(tptr[0], N, (int32_t*)qt_[c], (uint32_t*)at_[c], nullptr), "hot order (topology)");
    }
}

void GPUCache::CostModel(int, GPUNodeStorage* noder, GPUGraphStorage* graph, std::vector<uint64_t>& counters, int32_t train_step)
{
    const int64_t N = noder->TotalNodeNum();
    std::vector<int32_t> max_ids(kg_);
    uint64_t trans = counters.empty() ? 0 : counters[0];       // the reference passes Intel PCM's PCIe read count (Server.cu:84-108)
    for (int j = 0; j < kg_; j++) {
        lgn_ctx* ctx = LegionContextOfDevice(j);
        max_ids[j] = ctx ? lgn_max_ids(ctx, nullptr) : 0;
        if (counters.empty() && ctx) { int64_t tot[2]; if (!lgn_sampling_totals(ctx, nullptr, tot, 0)) trans += (uint64_t)(tot[0] + tot[1]); }
    }
    ABI_DIE(lgn_set_device(0), "SetGPUDevice");
    ABI_DIE(lgn_cost_model((uint32_t*)af_[0], (uint32_t*)at_[0], (int32_t*)qt_[0], graph->GetCSRNodeIndexCPU(), N, dim_, cache_memory_, kg_, trans,
                           max_ids.data(), train_step, &node_cap_, &edge_cap_, nullptr), "CostModel");
}

void GPUCache::FillUp(int, GPUNodeStorage* noder, GPUGraphStorage* graph)
{
    const int64_t N = noder->TotalNodeNum();
    const int kc = n_dev_ / kg_;
    shards_.assign(n_dev_, nullptr);
    for (int c = 0; c < kc; c++) {
        const int lead = c * kg_;
        std::vector<const float*> fshard(kg_);
        std::vector<const int64_t*> tptr(kg_);
        std::vector<const int32_t*> tidx(kg_);
        std::vector<int32_t*> fslot(kg_), tslot(kg_);
        for (int j = 0; j < kg_; j++) {
            ABI_DIE(lgn_set_device(lead + j), "SetGPUDevice");
            void *oq = qf_[c], *ot = qt_[c];
            if (j > 0) {
                ABI_DIE(lgn_device_alloc(&oq, N * 4), "alloc"); ABI_DIE(lgn_device_alloc(&ot, N * 4), "alloc");
                ABI_DIE(lgn_copy_d2d(oq, qf_[c], N * 4), "copy"); ABI_DIE(lgn_copy_d2d(ot, qt_[c], N * 4), "copy");
            }
            void *shard = nullptr, *fs = nullptr, *ts = nullptr, *tip = nullptr, *tix = nullptr;
            ABI_DIE(lgn_device_alloc(&shard, (int64_t)node_cap_ * dim_ * 4), "alloc shard");
            ABI_DIE(lgn_device_alloc(&fs, N * 4), "alloc"); ABI_DIE(lgn_device_alloc(&ts, N * 4), "alloc");
            ABI_DIE(lgn_fill_feature_shard((int32_t*)oq, N, node_cap_, kg_, j, noder->GetAllFloatAttr(), dim_, (float*)shard, nullptr), "FeatFillUp");
            ABI_DIE(lgn_place((int32_t*)oq, N, node_cap_, kg_, (int32_t*)fs, nullptr), "InitPair");
            ABI_DIE(lgn_place((int32_t*)ot, N, edge_cap_, kg_, (int32_t*)ts, nullptr), "InitIndexPair");
            ABI_DIE(lgn_device_alloc(&tip, (int64_t)(edge_cap_ + 1) * 8), "alloc");
            int64_t cnt = 0;
            ABI_DIE(lgn_fill_topo_shard((int32_t*)ot, N, edge_cap_, kg_, j, graph->GetCSRNodeIndexCPU(), graph->GetCSRNodeMatrixCPU(), (int64_t*)tip, nullptr, &cnt, nullptr), "GetNeighborCount");
            ABI_DIE(lgn_device_alloc(&tix, (cnt > 0 ? cnt : 1) * 4), "alloc");
            ABI_DIE(lgn_fill_topo_shard((int32_t*)ot, N, edge_cap_, kg_, j, graph->GetCSRNodeIndexCPU(), graph->GetCSRNodeMatrixCPU(), (int64_t*)tip, (int32_t*)tix, &cnt, nullptr), "TopoFillUp");
            ABI_DIE(lgn_device_synchronize(), "sync");
            fshard[j] = (float*)shard; fslot[j] = (int32_t*)fs; tslot[j] = (int32_t*)ts; tptr[j] = (int64_t*)tip; tidx[j] = (int32_t*)tix;
            shards_[lead + j] = shard;
            for (void* q : {shard, fs, ts, tip, tix}) owned_.push_back(q);
            if (j > 0) { lgn_device_free(oq); lgn_device_free(ot); }
        }
        for (int j = 0; j < kg_; j++) {
            lgn_ctx* ctx = LegionContextOfDevice(lead + j);
            ABI_DIE(lgn_set_part(ctx, j), "lgn_set_part");
            ABI_DIE(lgn_bind_feature_cache(ctx, kg_, fshard.data(), fslot[j], node_cap_), "bind feature cache");
            ABI_DIE(lgn_bind_topology_cache(ctx, kg_, tptr.data(), tidx.data(), tslot[j], edge_cap_), "bind topology cache");
        }
    }
    filled_ = true;
}

// ---- IPCEnv (CUDA_IPC_Service.cu:39-357) -----------------------------------------------------------------------
namespace {
class CUDAIPCEnv : public IPCEnv {
public:
    explicit CUDAIPCEnv(int32_t n) : n_(n) {}
    void Coordinate(BuildInfo* info) override
    {
        raw_batch_ = info->raw_batch_size; epochs_ = info->epoch; n_nodes_ = info->total_num_nodes; dim_ = info->float_attr_len;
        g_raw_batch = raw_batch_;
        ABI_DIE(lgn_coordinate(info->training_set_num.data(), info->validation_set_num.data(), info->testing_set_num.data(), n_, raw_batch_, epochs_, &steps_), "Coordinate");
        const int32_t st[3] = {steps_.train_step, steps_.valid_step, steps_.test_step};
        ABI_DIE(lgn_ipc_server_create(n_, st, &ipc_), "NewIPCEnv");
    }
    int32_t GetMaxStep() override { return steps_.max_step; }
    void InitializeSamplesBuffer(int32_t batch_size, int32_t, int32_t feature_dim, int32_t dev, int32_t) override
    {      // the slot buffers are the context's own; the handles of everything but the features go out now (:140-215)
        lgn_ctx* ctx = ensure_ctx(dev, n_nodes_, feature_dim > 0 ? feature_dim : dim_, batch_size);
        ABI_DIE(lgn_ipc_server_publish(ipc_, dev, ctx, 0), "InitializeSamplesBuffer");
    }
    void InitializeFeaturesBuffer(int32_t batch_size, int32_t, int32_t feature_dim, int32_t dev, int32_t) override
    {
        lgn_ctx* ctx = ensure_ctx(dev, n_nodes_, feature_dim > 0 ? feature_dim : dim_, batch_size);
        ABI_DIE(lgn_ipc_server_publish(ipc_, dev, ctx, 1), "InitializeFeaturesBuffer");
    }
    int32_t GetRawBatchsize() override { return raw_batch_; }
    int32_t GetLocalBatchId(int32_t g) override { return lgn_local_batch_id(&steps_, epochs_, g); }
    int32_t GetCurrentBatchsize(int32_t dev, int32_t mode) override
    {
        return mode == LGN_MODE_TRAIN ? raw_batch_ : (mode == LGN_MODE_VALID ? steps_.valid_batch[dev] : steps_.test_batch[dev]);
    }
    int32_t GetCurrentMode(int32_t g) override { return lgn_mode_of_step(&steps_, epochs_, g); }
    int32_t* GetIds(int32_t d, int32_t p) override { return view(d, p).ids; }
    float* GetFloatFeatures(int32_t d, int32_t p) override { return view(d, p).features; }
    int32_t* GetLabels(int32_t d, int32_t p) override { return view(d, p).labels; }
    int32_t* GetAggSrc(int32_t d, int32_t p) override { return view(d, p).agg_src; }
    int32_t* GetAggDst(int32_t d, int32_t p) override { return view(d, p).agg_dst; }
    int32_t* GetNodeCounter(int32_t d, int32_t p) override { return view(d, p).node_counter; }
    int32_t* GetEdgeCounter(int32_t d, int32_t p) override { return view(d, p).edge_counter; }
    void IPCPost(int32_t d, int32_t p) override { ABI_DIE(lgn_ipc_server_post(ipc_, d, p), "IPCPost"); }
    void IPCWait(int32_t d, int32_t p) override { ABI_DIE(lgn_ipc_server_wait(ipc_, d, p), "IPCWait"); }
    void Finalize() override { if (ipc_) lgn_ipc_server_destroy(ipc_); ipc_ = nullptr; }
    int32_t GetTrainStep() override { return steps_.train_step; }

private:
    lgn_batch_view view(int32_t d, int32_t p)
    {
        lgn_batch_view v;
        memset(&v, 0, sizeof(v));
        lgn_ctx* ctx = LegionContextOfDevice(d);
        if (ctx) lgn_batch_buffers(ctx, p, &v);
        return v;
    }
    int32_t n_ = 0, raw_batch_ = 0, epochs_ = 1, n_nodes_ = 0, dim_ = 0;
    lgn_steps steps_{};
    lgn_ipc_server* ipc_ = nullptr;
};

class PreSCCacheController : public CacheController {
public:
    void Initialize(int32_t dev_id, int32_t) override { dev_ = dev_id; }
    void Finalize() override {}
    int32_t MaxIdNum() override { lgn_ctx* c = LegionContextOfDevice(dev_); return c ? lgn_max_ids(c, nullptr) : 0; }

private:
    int32_t dev_ = 0;
};
}  // namespace

IPCEnv* NewIPCEnv(int32_t device_count) { return new CUDAIPCEnv(device_count); }
CacheController* NewPreSCCacheController(int32_t, int32_t) { return new PreSCCacheController(); }
