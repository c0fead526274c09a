// reference_abi.h -- the names libgpu_kernel.so exports in the reference, on top of the B200 C-ABI.
//
// The reference's kernel library is entered through C-linkage functions that take C++ objects
// (src/Kernels.cuh:24-93, src/GPU_Graph_Storage.cuh:38-39, src/GPU_Node_Storage.cuh:60-61) plus four C++ factories
// (src/CUDA_IPC_Service.h:33, src/GPUCache.cuh:66, src/Server.h:137-165).  This header restates those entry points
// with the same names, parameter lists and calling order, so that a runner written against the reference headers
// (Server.cu:169-335, Operator.cu:10-123) compiles and links against liblegion_b200.so.  The classes are stand-ins:
// they keep the reference's public method names and meanings, but their state is a handle on the per-GPU lgn_ctx
// (include/legion_b200.h) instead of the reference's bitmaps, cuckoo tables and pointer tables.
//
// Mapping (what each call does here):
//   batch_generator_kernel   lgn_batch_generate            (creates / binds the GPU's context on first use)
//   GPU_Random_Sampling      lgn_sample_hop(op_id/2 - 1)
//   get_feature_kernel       lgn_gather_segment((op_id-1)/2)
//   make_update_plan         lgn_finish_batch (hotness while the cache has not been filled yet)
//   update_cache             no-op, as in the reference (Kernels.cu:786-805)
//   GPUCache::CandidateSelection / CostModel / FillUp   lgn_hot_order / lgn_cost_model / lgn_place + lgn_fill_*
//   GPUMemoryPool::Set*      lgn_attach_buffers (the runner keeps ownership of the wire buffers)
#ifndef LEGION_B200_REFERENCE_ABI_H
#define LEGION_B200_REFERENCE_ABI_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

struct lgn_ctx;

// fields GPUGraphStore fills before Build (src/BuildInfo.h); the BaM / SSD members are out of scope
struct BuildInfo {
    std::vector<int32_t> shard_to_partition, shard_to_device;
    int32_t partition_count = 0;
    std::vector<int32_t> training_set_num;
    std::vector<std::vector<int32_t>> training_set_ids, training_labels;
    std::vector<int32_t> validation_set_num;
    std::vector<std::vector<int32_t>> validation_set_ids, validation_labels;
    std::vector<int32_t> testing_set_num;
    std::vector<std::vector<int32_t>> testing_set_ids, testing_labels;
    int32_t total_num_nodes = 0, int_attr_len = 0, float_attr_len = 0;
    int64_t* host_int_attrs = nullptr;
    float* host_float_attrs = nullptr;        // pinned + mapped (host_alloc_space)
    int64_t* csr_node_index = nullptr;        // pinned + mapped, N + 1 entries
    int32_t* csr_dst_node_ids = nullptr;      // pinned + mapped, E entries
    int64_t cache_edge_num = 0, total_edge_num = 0;
    int32_t epoch = 1, raw_batch_size = 0;
};

class GPUGraphStorage {
public:
    virtual ~GPUGraphStorage() = default;
    virtual void Build(BuildInfo* info) = 0;
    virtual void GraphCache(int32_t* QT, int32_t Ki, int32_t Kg, int32_t capacity) = 0;
    virtual void Finalize() = 0;
    virtual int32_t GetPartitionCount() const = 0;
    virtual int64_t* GetCSRNodeIndexCPU() const = 0;      // device alias of the host CSR
    virtual int32_t* GetCSRNodeMatrixCPU() const = 0;
    virtual int64_t Src_Size(int32_t part_id) const = 0;
    virtual int64_t Dst_Size(int32_t part_id) const = 0;
};
class GPUNodeStorage {
public:
    virtual ~GPUNodeStorage() = default;
    virtual void Build(BuildInfo* info) = 0;
    virtual void Finalize() = 0;
    virtual int32_t* GetTrainingSetIds(int32_t part_id) const = 0;
    virtual int32_t* GetValidationSetIds(int32_t part_id) const = 0;
    virtual int32_t* GetTestingSetIds(int32_t part_id) const = 0;
    virtual int32_t* GetTrainingLabels(int32_t part_id) const = 0;
    virtual int32_t* GetValidationLabels(int32_t part_id) const = 0;
    virtual int32_t* GetTestingLabels(int32_t part_id) const = 0;
    virtual int32_t TrainingSetSize(int32_t part_id) const = 0;
    virtual int32_t ValidationSetSize(int32_t part_id) const = 0;
    virtual int32_t TestingSetSize(int32_t part_id) const = 0;
    virtual int32_t TotalNodeNum() const = 0;
    virtual float* GetAllFloatAttr() const = 0;
    virtual int32_t GetFloatAttrLen() const = 0;
};

// src/GPUMemoryPool.cuh:7-208: per-runner registry of the batch buffers, double-buffered members indexed by the
// current pipe.  The sampling scratch (bitmap, position map, temporaries) has no counterpart: it lives in the lgn_ctx.
class GPUMemoryPool {
public:
    explicit GPUMemoryPool(int32_t pipeline_depth);
    int32_t GetIter() { return iter_; }
    int32_t GetCurrentMode() { return mode_; }
    int32_t GetOpId() { return op_id_; }
    float* GetFloatFeatures() { return float_features_[pipe_]; }
    int32_t* GetLabels() { return labels_[pipe_]; }
    int32_t* GetNodeCounter() { return node_counter_[pipe_]; }
    int32_t* GetEdgeCounter() { return edge_counter_[pipe_]; }
    int32_t* GetSampledIds() { return sampled_ids_[pipe_]; }
    int32_t* GetAggSrcOf() { return agg_src_off_[pipe_]; }
    int32_t* GetAggDstOf() { return agg_dst_off_[pipe_]; }
    int32_t* GetAggSrcId();            // raw ids of the current slot (lane-private here; shared between pipes in the reference)
    int32_t* GetAggDstId();
    void SetFloatFeatures(float* p, int32_t pipe) { float_features_[pipe] = p; dirty_ = true; }
    void SetLabels(int32_t* p, int32_t pipe) { labels_[pipe] = p; dirty_ = true; }
    void SetNodeCounter(int32_t* p, int32_t pipe) { node_counter_[pipe] = p; dirty_ = true; }
    void SetEdgeCounter(int32_t* p, int32_t pipe) { edge_counter_[pipe] = p; dirty_ = true; }
    void SetSampledIds(int32_t* p, int32_t pipe) { sampled_ids_[pipe] = p; dirty_ = true; }
    void SetAggSrcOf(int32_t* p, int32_t pipe) { agg_src_off_[pipe] = p; dirty_ = true; }
    void SetAggDstOf(int32_t* p, int32_t pipe) { agg_dst_off_[pipe] = p; dirty_ = true; }
    // sizes of the buffers registered above (the reference passes them to cudaMalloc only): entries of the id / edge
    // arrays and rows of the feature buffer.  Defaults: the context's worst case.
    void SetBufferSizes(int64_t num_ids, int64_t feature_rows) { num_ids_ = num_ids; feature_rows_ = feature_rows; dirty_ = true; }
    // accepted and ignored: scratch the reference runner allocates for its own kernels (Server.cu:233-283)
    void SetCacheSearchBuffer(int32_t*) {}
    void SetAccessedMap(uint32_t*) {}
    void SetPositionMap(int32_t*) {}
    void SetAggSrcId(int32_t*) {}
    void SetAggDstId(int32_t*) {}
    void SetTmpSrcOf(int32_t*) {}
    void SetTmpDstOf(int32_t*) {}
    void SetTempStorage(void*) {}
    void SetTmpPartIdx(char*) {}
    void SetTmpPartOff(int32_t*) {}
    void SetOpId(int32_t op_id) { op_id_ = op_id; }
    void SetCurrentPipe(int32_t pipe) { pipe_ = pipe; }
    void SetCurrentMode(int32_t mode) { mode_ = mode; }
    void SetIter(int32_t iter) { iter_ = iter; }
    void Finalize() {}
    // --- this side only
    lgn_ctx* ctx = nullptr;
    int32_t pipe() const { return pipe_; }
    void Sync();                       // push Set* changes into the context (lgn_attach_buffers)

private:
    static const int kDepth = 2;
    int32_t iter_ = 0, mode_ = 0, op_id_ = 0, pipe_ = 0, depth_ = 2;
    bool dirty_ = false;
    int64_t num_ids_ = 0, feature_rows_ = 0;
    float* float_features_[kDepth] = {nullptr, nullptr};
    int32_t *labels_[kDepth] = {}, *node_counter_[kDepth] = {}, *edge_counter_[kDepth] = {}, *sampled_ids_[kDepth] = {};
    int32_t *agg_src_off_[kDepth] = {}, *agg_dst_off_[kDepth] = {};
};

// src/GPUCache.cuh:68-160: presampling statistics, candidate selection, cost model, fill-up
class GPUCache {
public:
    void Initialize(int64_t cache_memory, int32_t int_attr_len, int32_t float_attr_len, int32_t train_step, int32_t device_count);
    void InitializeCacheController(int32_t dev_id, int32_t total_num_nodes);
    void Finalize(int32_t dev_id);
    int32_t NodeCapacity(int32_t dev_id);
    void CandidateSelection(int cache_agg_mode, GPUNodeStorage* noder, GPUGraphStorage* graph);
    void CostModel(int cache_agg_mode, GPUNodeStorage* noder, GPUGraphStorage* graph, std::vector<uint64_t>& counters, int32_t train_step);
    void FillUp(int cache_agg_mode, GPUNodeStorage* noder, GPUGraphStorage* graph);
    float* Float_Feature_Cache(int32_t dev_id);
    int32_t MaxIdNum(int32_t dev_id);
    bool filled() const { return filled_; }

private:
    int64_t cache_memory_ = 0;
    int32_t dim_ = 0, train_step_ = 0, n_dev_ = 0, kg_ = 1;
    int32_t node_cap_ = 0, edge_cap_ = 0;
    bool filled_ = false;
    std::vector<void*> qf_, qt_, af_, at_;      // per clique leader: hot orders and sorted counts (device)
    std::vector<void*> shards_;
    std::vector<void*> owned_;
};

extern "C" {
void* d_alloc_space(int64_t num_bytes);
void* d_alloc_space_managed(unsigned int num_bytes);
void d_copy_2_h(void* h_ptr, void* d_ptr, unsigned int num_bytes);
void d_free_space(void* d_ptr);
void SetGPUDevice(int32_t shard_id);
int32_t GetGPUDevice();
void* host_alloc_space(unsigned int num_bytes);      // returns the DEVICE alias of pinned, mapped host memory (Kernels.cu:57-64)
void batch_generator_kernel(cudaStream_t strm_hdl, GPUNodeStorage* noder, GPUCache* cache, GPUMemoryPool* memorypool,
                            int32_t batch_size, int32_t counter, int32_t part_id, int32_t dev_id, int32_t mode);
void GPU_Random_Sampling(cudaStream_t strm_hdl, GPUGraphStorage* graph, GPUCache* cache, GPUMemoryPool* memorypool,
                         int32_t count, int32_t op_id, bool is_presc);
void get_feature_kernel(cudaStream_t strm_hdl, GPUCache* cache, GPUNodeStorage* noder, GPUMemoryPool* memorypool,
                        int32_t dev_id, int32_t op_id, bool in_memory);
void make_update_plan(cudaStream_t strm_hdl, GPUGraphStorage* graph, GPUCache* cache, GPUMemoryPool* memorypool,
                      int32_t dev_id, int32_t mode);
void update_cache(cudaStream_t strm_hdl, GPUCache* cache, GPUNodeStorage* noder, GPUMemoryPool* memorypool,
                  int32_t dev_id, int32_t mode);
GPUGraphStorage* NewGPUMemoryGraphStorage();
GPUNodeStorage* NewGPUMemoryNodeStorage();
}

// the wire-format side (src/CUDA_IPC_Service.h:6-35)
class IPCEnv {
public:
    virtual ~IPCEnv() = default;
    virtual void Coordinate(BuildInfo* info) = 0;
    virtual int32_t GetMaxStep() = 0;
    virtual void InitializeSamplesBuffer(int32_t batch_size, int32_t num_ids, int32_t feature_dim, int32_t device_id, int32_t pipeline_depth) = 0;
    virtual void InitializeFeaturesBuffer(int32_t batch_size, int32_t num_ids, int32_t feature_dim, int32_t device_id, int32_t pipeline_depth) = 0;
    virtual int32_t GetRawBatchsize() = 0;
    virtual int32_t GetLocalBatchId(int32_t global_batch_id) = 0;
    virtual int32_t GetCurrentBatchsize(int32_t dev_id, int32_t current_mode) = 0;
    virtual int32_t GetCurrentMode(int32_t global_batch_id) = 0;
    virtual int32_t* GetIds(int32_t dev_id, int32_t current_pipe) = 0;
    virtual float* GetFloatFeatures(int32_t dev_id, int32_t current_pipe) = 0;
    virtual int32_t* GetLabels(int32_t dev_id, int32_t current_pipe) = 0;
    virtual int32_t* GetAggSrc(int32_t dev_id, int32_t current_pipe) = 0;
    virtual int32_t* GetAggDst(int32_t dev_id, int32_t current_pipe) = 0;
    virtual int32_t* GetNodeCounter(int32_t dev_id, int32_t current_pipe) = 0;
    virtual int32_t* GetEdgeCounter(int32_t dev_id, int32_t current_pipe) = 0;
    virtual void IPCPost(int32_t dev_id, int32_t current_pipe) = 0;
    virtual void IPCWait(int32_t dev_id, int32_t current_pipe) = 0;
    virtual void Finalize() = 0;
    virtual int32_t GetTrainStep() = 0;
};
IPCEnv* NewIPCEnv(int32_t device_count);

// src/GPUCache.cuh:9-64.  The lookups of the reference controller (two cuckoo tables) are direct-mapped slot tables
// inside the lgn_ctx; the controller object only answers the queries a runner makes.
class CacheController {
public:
    virtual ~CacheController() = default;
    virtual void Initialize(int32_t dev_id, int32_t total_num_nodes) = 0;
    virtual void Finalize() = 0;
    virtual int32_t MaxIdNum() = 0;
};
CacheController* NewPreSCCacheController(int32_t train_step, int32_t device_count);

// every context created through this layer (for tests / shutdown)
lgn_ctx* LegionContextOfDevice(int32_t dev_id);
void LegionReleaseContexts();

#endif  // LEGION_B200_REFERENCE_ABI_H
