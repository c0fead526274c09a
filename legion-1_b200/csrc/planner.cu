// planner.cu -- cache planning after presampling: hot order, placement, shard fill,
// cost model.  One-off work per server start (not on the per-batch path).
//
// Replaces GPUCache::CandidateSelection / CostModel / FillUp (GPUCache.cu:578-826),
// InitPair / InitIndexPair / InitOffsetPair (GPUCache.cu:88-108), the bght::bcht cuckoo
// tables (src/include/bcht.hpp) and GPUMemoryGraphStorage::GraphCache
// (GPU_Memory_Graph_Storage.cu:14-35, 98-133).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "context.h"

extern thread_local char g_lgn_cuda_err[256];
int lgn_cuda_fail(cudaError_t e, const char* what);
#define CK(x)                                                 \
    do {                                                      \
        cudaError_t e_ = (x);                                 \
        if (e_ != cudaSuccess) return lgn_cuda_fail(e_, #x);  \
    } while (0)

namespace lgn {

__global__ void k_iota(int32_t* p, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

// slot_of[id] = (i % kg) * cap + i / kg for hot rank i (InitPair, GPUCache.cu:103-108).
// A direct-mapped table: one 4-byte probe per lookup instead of a cuckoo bucket walk.
__global__ void k_place(const int32_t* __restrict__ order, long long n, long long cap, int kg, long long n_repl, int my_part,
                        int32_t* __restrict__ slot_of)
{
    const long long lim = n_repl + (cap - n_repl) * kg;   // ranks that have a row somewhere in the clique
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int32_t id = order[i];
        int32_t slot = -1;
        if (i < n_repl) slot = (int32_t)(my_part * cap + i);                 // replicated: every GPU serves it locally
        else if (i < lim) { const long long k = i - n_repl; slot = (int32_t)((k % kg) * cap + n_repl + k / kg); }
        slot_of[id] = slot;
    }
}

// neighbour counts of shard j's rows (GetNeighborCount, GPU_Memory_Graph_Storage.cu:14-20)
__global__ void k_shard_degrees(const int32_t* __restrict__ order, long long n, long long cap, int kg, int j,
                                const int64_t* __restrict__ indptr, int64_t* __restrict__ deg)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cap; t += (long long)gridDim.x * blockDim.x) {
        const long long rank = t * kg + j;
        long long d = 0;
        if (rank < n) { const int32_t id = order[rank]; d = indptr[id + 1] - indptr[id]; }
        deg[t] = d;
    }
}

// warp-per-node adjacency copy (TopoFillUp, GPU_Memory_Graph_Storage.cu:22-35 runs one
// thread per node with a serial inner loop over UVA memory)
__global__ void k_shard_adjacency(const int32_t* __restrict__ order, long long n, long long cap, int kg, int j,
                                  const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                  const int64_t* __restrict__ indptr_out, int32_t* __restrict__ indices_out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = warp; t < cap; t += n_warps) {
        const long long rank = t * kg + j;
        if (rank >= n) continue;
        const int32_t id = order[rank];
        const long long s = indptr[id], d = indptr[id + 1] - s, o = indptr_out[t];
        for (long long k = lane; k < d; k += 32) indices_out[o + k] = indices[s + k];
    }
}

// adjacency bytes per node in topology order (GetEdgeMem, GPUCache.cu:35-41)
__global__ void k_edge_mem(const int32_t* __restrict__ qt, long long n, const int64_t* __restrict__ indptr,
                           unsigned long long* __restrict__ mem)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int32_t id = qt[i];
        mem[i] = 8ull + 4ull * (unsigned long long)(indptr[id + 1] - indptr[id]);
    }
}

__global__ void k_widen(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, long long n)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

}  // namespace lgn

using namespace lgn;

extern "C" int lgn_hot_order(const uint32_t* counts, int64_t n, int32_t* order, uint32_t* sorted_counts, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!counts || !order || n <= 0 || n > 0x7fffffffLL) return LGN_E_ARG;
    int32_t* iota = nullptr;
    uint32_t* keys_out = sorted_counts;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    CK(cudaMalloc(&iota, n * sizeof(int32_t)));
    if (!keys_out) CK(cudaMalloc(&keys_out, n * sizeof(uint32_t)));
    k_iota<<<1024, 256, 0, s>>>(iota, n);
    // stable LSD radix sort, descending keys, payload = iota  ==> (count desc, id asc):
    // the order thrust::sort_by_key(greater) yields in the reference (GPUCache.cu:630-631)
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, counts, keys_out, iota, order, (int)n, 0, 32, s));
    CK(cudaMalloc(&tmp, tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, counts, keys_out, iota, order, (int)n, 0, 32, s));
    CK(cudaStreamSynchronize(s));
    cudaFree(tmp);
    cudaFree(iota);
    if (!sorted_counts) cudaFree(keys_out);
    return LGN_OK;
}

extern "C" int lgn_place_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int64_t n_repl, int32_t my_part,
                                int32_t* slot_of, void* stream)
{
    if (!order || !slot_of || n <= 0 || kg <= 0 || kg > LGN_MAX_PARTS || cap < 0) return LGN_E_ARG;
    if (n_repl < 0 || n_repl > cap || my_part < 0 || my_part >= kg) return LGN_E_ARG;
    if (cap * kg > 0x7fffffffLL) return LGN_E_ARG;   // reference overflows int32 here (GPUCache.cu:315)
    k_place<<<1024, 256, 0, (cudaStream_t)stream>>>(order, n, cap, kg, n_repl, my_part, slot_of);
    CK(cudaGetLastError());
    return LGN_OK;
}

extern "C" int lgn_place(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t* slot_of, void* stream)
{
    return lgn_place_hybrid(order, n, cap, kg, 0, 0, slot_of, stream);
}

extern "C" int lgn_fill_feature_shard_hybrid(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j, int64_t n_repl,
                                             const float* features, int32_t dim, float* shard, void* stream)
{
    if (!order || !features || !shard || kg <= 0 || j < 0 || j >= kg || dim <= 0 || n_repl < 0 || n_repl > cap) return LGN_E_ARG;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    launch_row_copy(order, n, cap, kg, j, n_repl, features, dim, shard, n_sm, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return LGN_OK;
}

extern "C" int lgn_fill_feature_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                                      const float* features, int32_t dim, float* shard, void* stream)
{
    return lgn_fill_feature_shard_hybrid(order, n, cap, kg, j, 0, features, dim, shard, stream);
}

extern "C" int lgn_fill_topo_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                                   const int64_t* indptr, const int32_t* indices, int64_t* indptr_out,
                                   int32_t* indices_out, int64_t* n_indices, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!order || !indptr || !indptr_out || kg <= 0 || j < 0 || j >= kg || cap <= 0) return LGN_E_ARG;
    if (!indices_out) {
        int64_t* deg = nullptr;
        void* tmp = nullptr;
        size_t tmp_bytes = 0;
        CK(cudaMalloc(&deg, cap * sizeof(int64_t)));
        k_shard_degrees<<<1024, 256, 0, s>>>(order, n, cap, kg, j, indptr, deg);
        CK(cudaMemsetAsync(indptr_out, 0, sizeof(int64_t), s));
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, deg, indptr_out + 1, (int)cap, s));
        CK(cudaMalloc(&tmp, tmp_bytes));
        CK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, deg, indptr_out + 1, (int)cap, s));
        int64_t total = 0;
        CK(cudaMemcpyAsync(&total, indptr_out + cap, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        cudaFree(tmp);
        cudaFree(deg);
        if (n_indices) *n_indices = total;
        return LGN_OK;
    }
    if (!indices) return LGN_E_ARG;
    k_shard_adjacency<<<1024, 256, 0, s>>>(order, n, cap, kg, j, indptr, indices, indptr_out, indices_out);
    CK(cudaGetLastError());
    return LGN_OK;
}

// CostModel (GPUCache.cu:661-767): prefix sums on the device, the 100-step sweep on the
// host with the reference's arithmetic (float tables, double ratios).
extern "C" int lgn_cost_model(const uint32_t* af, const uint32_t* at, const int32_t* qt, const int64_t* indptr, int64_t n,
                              int32_t dim, int64_t cache_memory, int32_t kg, uint64_t topo_trans, const int32_t* max_ids,
                              int32_t train_step, int32_t* node_capacity, int32_t* edge_capacity)
{
    if (!af || !at || !qt || !indptr || n <= 0 || kg <= 0 || !max_ids || !node_capacity || !edge_capacity) return LGN_E_ARG;
    unsigned long long *w = nullptr, *node_prefix = nullptr, *edge_prefix = nullptr, *mem_prefix = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    CK(cudaMalloc(&w, n * 8)); CK(cudaMalloc(&node_prefix, n * 8)); CK(cudaMalloc(&edge_prefix, n * 8)); CK(cudaMalloc(&mem_prefix, n * 8));
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, w, node_prefix, (int)n));
    CK(cudaMalloc(&tmp, tmp_bytes));
    k_widen<<<1024, 256>>>(af, w, n);
    CK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, w, node_prefix, (int)n));
    k_widen<<<1024, 256>>>(at, w, n);
    CK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, w, edge_prefix, (int)n));
    k_edge_mem<<<1024, 256>>>(qt, n, indptr, w);
    CK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, w, mem_prefix, (int)n));
    std::vector<unsigned long long> h_node(n), h_edge(n), h_mem(n);
    CK(cudaMemcpy(h_node.data(), node_prefix, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_edge.data(), edge_prefix, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_mem.data(), mem_prefix, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(w); cudaFree(node_prefix); cudaFree(edge_prefix); cudaFree(mem_prefix); cudaFree(tmp);

    const int max_payload = 64;                                                   // CLS, GPUCache.cu:31
    int64_t memory_step = (int64_t)((double)(cache_memory * kg) * 0.01);          // :674
    if (memory_step < 1) memory_step = 1;
    uint64_t feat_trans = 0;
    for (int i = 0; i < kg; i++)                                                  // :677-679
        feat_trans += (uint64_t)(((((int64_t)max_ids[i] * train_step) * dim) * (int64_t)sizeof(float)) / max_payload);
    const int64_t total_mem = cache_memory * kg;
    const int64_t steps = (total_mem - 1) / memory_step + 1;
    std::vector<float> t_topo(steps + 1, 0.f), t_feat(steps + 1, 0.f), c_topo(steps + 1, 0.f), c_feat(steps + 1, 0.f), t_total(steps + 1, 0.f);
    int64_t cs = 0;
    for (int64_t cur = 0; cur < total_mem; cur += memory_step) {                  // :723-753
        int32_t nf, nt;
        if ((uint64_t)cur > (uint64_t)n * dim * sizeof(float)) nf = (int32_t)n;
        else nf = (int32_t)((uint64_t)(cs + 1) * ((uint64_t)memory_step / (dim * sizeof(float))));
        if ((uint64_t)cur > h_mem[n - 1]) nt = (int32_t)n;
        else nt = (int32_t)(std::lower_bound(h_mem.begin(), h_mem.end(), (unsigned long long)cur) - h_mem.begin());
        if (nt < n) {
            const unsigned long long ep = nt > 0 ? h_edge[nt - 1] : 0;            // reference reads [-1] at step 0 (unused)
            t_topo[cs] = (float)((double)topo_trans * 1.0 / (double)h_edge[n - 1] * (double)ep);
            c_topo[cs] = (float)(nt / kg);
        }
        if (nf < n) {
            const unsigned long long np = nf > 0 ? h_node[nf - 1] : 0;
            t_feat[cs] = (float)((double)feat_trans * 1.0 / (double)h_node[n - 1] * (double)np);
            c_feat[cs] = (float)(nf / kg);
        }
        cs++;
    }
    for (int64_t sidx = 1; sidx < steps; sidx++) t_total[sidx] = t_topo[sidx] + t_feat[steps - 1 - sidx];   // :755-760
    const int64_t best = std::max_element(t_total.begin(), t_total.end()) - t_total.begin();
    *node_capacity = (int32_t)(c_feat[steps - 1 - best] + 1);                     // :763-764
    *edge_capacity = (int32_t)(c_topo[best] + 1);
    return LGN_OK;
}
