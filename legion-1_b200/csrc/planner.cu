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

// ---- compact placement map (include/legion_b200.h: lgn_place_compact) ----
// record r = 8 words: [0] replicated nodes before node r*96, [1] partitioned nodes before it, [2..4] replicated bits,
// [5..7] partitioned bits of nodes r*96 .. r*96+95
__global__ void k_cmap_mark(const int32_t* __restrict__ order, long long n_repl, long long n_cached, uint32_t* __restrict__ w)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cached; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t id = (uint32_t)order[i];
        const uint32_t rec = id / (uint32_t)LGN_CMAP_NODES, j = id - rec * (uint32_t)LGN_CMAP_NODES;
        atomicOr(&w[(size_t)rec * 8 + (i < n_repl ? 2 : 5) + (j >> 5)], 1u << (j & 31u));
    }
}
__global__ void k_cmap_count(const uint32_t* __restrict__ w, long long n_rec, uint32_t* __restrict__ cnt_r, uint32_t* __restrict__ cnt_p)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (long long)gridDim.x * blockDim.x) {
        const uint32_t* q = w + (size_t)r * 8;
        cnt_r[r] = __popc(q[2]) + __popc(q[3]) + __popc(q[4]);
        cnt_p[r] = __popc(q[5]) + __popc(q[6]) + __popc(q[7]);
    }
}
__global__ void k_cmap_prefix(uint32_t* __restrict__ w, long long n_rec, const uint32_t* __restrict__ pre_r, const uint32_t* __restrict__ pre_p)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += (long long)gridDim.x * blockDim.x) {
        w[(size_t)r * 8] = pre_r[r];
        w[(size_t)r * 8 + 1] = pre_p[r];
    }
}
// shard j of the compact placement: one warp per 32 consecutive nodes, lane l classifies node base+l, the warp copies the
// rows that belong to this shard.  Reads of the feature matrix and writes of the shard both advance in node-id order.
__global__ void __launch_bounds__(256) k_cmap_fill(const uint32_t* __restrict__ w, long long n, long long n_repl, int kg, int part,
                                                   const float* __restrict__ src, int dim, float* __restrict__ dst, long long cap)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool vec = (dim & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0;
    for (long long n0 = warp * 32; n0 < n; n0 += n_warps * 32) {
        const long long nid = n0 + lane;
        long long row = -1;
        if (nid < n) {
            const uint32_t rec = (uint32_t)nid / (uint32_t)LGN_CMAP_NODES, j = (uint32_t)nid - rec * (uint32_t)LGN_CMAP_NODES;
            const uint32_t* q = w + (size_t)rec * 8;
            const uint32_t wi = j >> 5, bit = j & 31u, below = (1u << bit) - 1u;
            if ((q[2 + wi] >> bit) & 1u) {
                uint32_t c = q[0] + __popc(q[2 + wi] & below);
                for (uint32_t k = 0; k < wi; k++) c += __popc(q[2 + k]);
                row = c;
            } else if ((q[5 + wi] >> bit) & 1u) {
                uint32_t c = q[1] + __popc(q[5 + wi] & below);
                for (uint32_t k = 0; k < wi; k++) c += __popc(q[5 + k]);
                if ((int)(c % (uint32_t)kg) == part) row = n_repl + c / (uint32_t)kg;
            }
            if (row >= cap) row = -1;
        }
        uint32_t todo = __ballot_sync(0xffffffffu, row >= 0);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long r = __shfl_sync(0xffffffffu, row, l);
            const float* s = src + (n0 + l) * dim;
            float* d = dst + r * dim;
            if (vec) {
                for (int k = lane; k < (dim >> 2); k += 32) reinterpret_cast<uint4*>(d)[k] = ld_nc_v4(reinterpret_cast<const uint4*>(s) + k);
            } else {
                for (int k = lane; k < dim; k += 32) d[k] = s[k];
            }
        }
    }
}

}  // namespace lgn

using namespace lgn;

// temporary device buffer released on every exit path
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() const { return (T*)p; }
};

extern "C" int lgn_hot_order(const uint32_t* counts, int64_t n, int32_t* order, uint32_t* sorted_counts, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!counts || !order || n <= 0 || n > 0x7fffffffLL) return LGN_E_ARG;
    DevBuf iota, keys, tmp;
    size_t tmp_bytes = 0;
    CK(iota.alloc(n * sizeof(int32_t)));
    uint32_t* keys_out = sorted_counts;
    if (!keys_out) { CK(keys.alloc(n * sizeof(uint32_t))); keys_out = keys.as<uint32_t>(); }
    k_iota<<<1024, 256, 0, s>>>(iota.as<int32_t>(), n);
    // stable LSD radix sort, descending keys, payload = iota  ==> (count desc, id asc):
    // the order thrust::sort_by_key(greater) yields in the reference (GPUCache.cu:630-631)
    CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, counts, keys_out, iota.as<int32_t>(), order, (int)n, 0, 32, s));
    CK(tmp.alloc(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, counts, keys_out, iota.as<int32_t>(), order, (int)n, 0, 32, s));
    CK(cudaStreamSynchronize(s));
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

extern "C" int64_t lgn_cmap_bytes(int64_t n_nodes)
{
    if (n_nodes <= 0) return 0;
    return ((n_nodes + LGN_CMAP_NODES - 1) / LGN_CMAP_NODES) * 32;
}

extern "C" int lgn_place_compact(const int32_t* order, int64_t n, int64_t n_repl, int64_t n_part, void* cmap, void* stream)
{
    if (!order || !cmap || n <= 0 || n > 0x7fffffffLL || n_repl < 0 || n_part < 0 || n_repl + n_part > n) return LGN_E_ARG;
    if ((uintptr_t)cmap & 15) return LGN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_rec = (n + LGN_CMAP_NODES - 1) / LGN_CMAP_NODES;
    uint32_t* w = (uint32_t*)cmap;
    DevBuf cnt_r, cnt_p, pre_r, pre_p, tmp;
    size_t tmp_bytes = 0;
    CK(cnt_r.alloc(n_rec * 4)); CK(cnt_p.alloc(n_rec * 4)); CK(pre_r.alloc(n_rec * 4)); CK(pre_p.alloc(n_rec * 4));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt_r.as<uint32_t>(), pre_r.as<uint32_t>(), (int)n_rec, s));
    CK(tmp.alloc(tmp_bytes));
    CK(cudaMemsetAsync(w, 0, (size_t)n_rec * 32, s));
    if (n_repl + n_part > 0) k_cmap_mark<<<1024, 256, 0, s>>>(order, n_repl, n_repl + n_part, w);
    k_cmap_count<<<1024, 256, 0, s>>>(w, n_rec, cnt_r.as<uint32_t>(), cnt_p.as<uint32_t>());
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt_r.as<uint32_t>(), pre_r.as<uint32_t>(), (int)n_rec, s));
    CK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt_p.as<uint32_t>(), pre_p.as<uint32_t>(), (int)n_rec, s));
    k_cmap_prefix<<<1024, 256, 0, s>>>(w, n_rec, pre_r.as<uint32_t>(), pre_p.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));          // the temporaries are released on return
    return LGN_OK;
}

extern "C" int lgn_fill_feature_shard_compact(const void* cmap, int64_t n, int64_t n_repl, int32_t kg, int32_t j, const float* features,
                                              int32_t dim, float* shard, int64_t cap, void* stream)
{
    if (!cmap || !features || !shard || n <= 0 || kg <= 0 || kg > LGN_MAX_PARTS || j < 0 || j >= kg || dim <= 0 || n_repl < 0 || cap < n_repl)
        return LGN_E_ARG;
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    k_cmap_fill<<<n_sm * 8, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)cmap, n, n_repl, kg, j, features, dim, shard, cap);
    CK(cudaGetLastError());
    return LGN_OK;
}

extern "C" int lgn_fill_topo_shard(const int32_t* order, int64_t n, int64_t cap, int32_t kg, int32_t j,
                                   const int64_t* indptr, const int32_t* indices, int64_t* indptr_out,
                                   int32_t* indices_out, int64_t* n_indices, void* stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (!order || !indptr || !indptr_out || kg <= 0 || j < 0 || j >= kg || cap <= 0) return LGN_E_ARG;
    if (!indices_out) {
        DevBuf deg, tmp;
        size_t tmp_bytes = 0;
        CK(deg.alloc(cap * sizeof(int64_t)));
        k_shard_degrees<<<1024, 256, 0, s>>>(order, n, cap, kg, j, indptr, deg.as<int64_t>());
        CK(cudaMemsetAsync(indptr_out, 0, sizeof(int64_t), s));
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, deg.as<int64_t>(), indptr_out + 1, (int)cap, s));
        CK(tmp.alloc(tmp_bytes));
        CK(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, deg.as<int64_t>(), indptr_out + 1, (int)cap, s));
        int64_t total = 0;
        CK(cudaMemcpyAsync(&total, indptr_out + cap, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (n_indices) *n_indices = total;
        return LGN_OK;
    }
    if (!indices) return LGN_E_ARG;
    k_shard_adjacency<<<1024, 256, 0, s>>>(order, n, cap, kg, j, indptr, indices, indptr_out, indices_out);
    CK(cudaGetLastError());
    return LGN_OK;
}

namespace lgn {
// one sweep point of the cost model: how many nodes' adjacency lists / feature rows fit in `mem` bytes and how much
// presampled hotness they cover.  The three prefix arrays stay on the device (the reference copies all of them to the
// host, 3 x 8N bytes); one thread per sweep point does the binary search.
struct SweepPoint { long long nt; unsigned long long edge_hot, node_hot; };
__global__ void k_sweep(const unsigned long long* __restrict__ node_prefix, const unsigned long long* __restrict__ edge_prefix,
                        const unsigned long long* __restrict__ mem_prefix, long long n, long long memory_step, long long steps,
                        const long long* __restrict__ nf_of_step, SweepPoint* __restrict__ out)
{
    const long long cs = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cs >= steps) return;
    const unsigned long long cur = (unsigned long long)(cs * memory_step);
    long long nt;
    if (cur > mem_prefix[n - 1]) nt = n;
    else {                                     // first index whose cumulative adjacency bytes reach `cur`
        long long lo = 0, hi = n;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (mem_prefix[mid] < cur) lo = mid + 1; else hi = mid; }
        nt = lo;
    }
    const long long nf = nf_of_step[cs];
    SweepPoint sp;
    sp.nt = nt;
    sp.edge_hot = (nt > 0 && nt <= n) ? edge_prefix[nt - 1] : 0;
    sp.node_hot = (nf > 0 && nf <= n) ? node_prefix[nf - 1] : 0;
    out[cs] = sp;
}
}  // namespace lgn

// CostModel (GPUCache.cu:661-767).  Capacities must equal the reference's bit for bit (tests/test_reference_ab.py), so
// the table arithmetic keeps its float / double conversions; the data movement does not follow it: prefix sums and the
// per-step lookups run on the device, 101 sweep points come back instead of three N-element arrays.
extern "C" int lgn_cost_model(const uint32_t* af, const uint32_t* at, const int32_t* qt, const int64_t* indptr, int64_t n,
                              int32_t dim, int64_t cache_memory, int32_t kg, uint64_t topo_trans, const int32_t* max_ids,
                              int32_t train_step, int32_t* node_capacity, int32_t* edge_capacity, void* stream)
{
    if (!af || !at || !qt || !indptr || n <= 0 || kg <= 0 || dim <= 0 || cache_memory <= 0 || !max_ids || !node_capacity || !edge_capacity)
        return LGN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total_mem = cache_memory * kg;
    int64_t memory_step = (int64_t)((double)total_mem * 0.01);                    // 1 % of the clique's budget (:674)
    if (memory_step < 1) memory_step = 1;
    const int64_t steps = (total_mem - 1) / memory_step + 1;
    const uint64_t row = (uint64_t)dim * sizeof(float);
    std::vector<long long> nf_of_step(steps);
    for (int64_t cs = 0; cs < steps; cs++) {                                      // feature rows that fit at sweep point cs (:728-731)
        const uint64_t cur = (uint64_t)(cs * memory_step);
        nf_of_step[cs] = cur > (uint64_t)n * row ? (long long)n : (long long)(int32_t)((uint64_t)(cs + 1) * ((uint64_t)memory_step / row));
    }
    DevBuf w, node_prefix, edge_prefix, mem_prefix, tmp, d_nf, d_out;
    size_t tmp_bytes = 0;
    CK(w.alloc(n * 8)); CK(node_prefix.alloc(n * 8)); CK(edge_prefix.alloc(n * 8)); CK(mem_prefix.alloc(n * 8));
    CK(d_nf.alloc(steps * sizeof(long long))); CK(d_out.alloc(steps * sizeof(SweepPoint)));
    auto* W = w.as<unsigned long long>();
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, W, node_prefix.as<unsigned long long>(), (int)n, s));
    CK(tmp.alloc(tmp_bytes));
    k_widen<<<1024, 256, 0, s>>>(af, W, n);
    CK(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, W, node_prefix.as<unsigned long long>(), (int)n, s));
    k_widen<<<1024, 256, 0, s>>>(at, W, n);
    CK(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, W, edge_prefix.as<unsigned long long>(), (int)n, s));
    k_edge_mem<<<1024, 256, 0, s>>>(qt, n, indptr, W);
    CK(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, W, mem_prefix.as<unsigned long long>(), (int)n, s));
    CK(cudaMemcpyAsync(d_nf.p, nf_of_step.data(), steps * sizeof(long long), cudaMemcpyHostToDevice, s));
    k_sweep<<<(unsigned)((steps + 127) / 128), 128, 0, s>>>(node_prefix.as<unsigned long long>(), edge_prefix.as<unsigned long long>(),
                                                           mem_prefix.as<unsigned long long>(), n, memory_step, steps, d_nf.as<long long>(),
                                                           d_out.as<SweepPoint>());
    std::vector<SweepPoint> pts(steps);
    unsigned long long node_total = 0, edge_total = 0;
    CK(cudaMemcpyAsync(pts.data(), d_out.p, steps * sizeof(SweepPoint), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&node_total, node_prefix.as<unsigned long long>() + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&edge_total, edge_prefix.as<unsigned long long>() + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));

    uint64_t feat_trans = 0;                                                      // 64-byte transactions of one epoch's gathers (:677-679)
    for (int i = 0; i < kg; i++) feat_trans += (uint64_t)(((((int64_t)max_ids[i] * train_step) * dim) * (int64_t)sizeof(float)) / 64);
    // saved transactions per sweep point: topology gets the first cs per cent of the budget, features the rest (:723-760)
    std::vector<float> saved_topo(steps + 1, 0.f), saved_feat(steps + 1, 0.f), cap_topo(steps + 1, 0.f), cap_feat(steps + 1, 0.f);
    for (int64_t cs = 0; cs < steps; cs++) {
        const long long nt = pts[cs].nt, nf = nf_of_step[cs];
        if (nt < n) {      // a tier that fits completely keeps gain 0 (the reference's quirk, :744-751)
            saved_topo[cs] = (float)((double)topo_trans * 1.0 / (double)edge_total * (double)pts[cs].edge_hot);
            cap_topo[cs] = (float)(nt / kg);
        }
        if (nf < n) {
            saved_feat[cs] = (float)((double)feat_trans * 1.0 / (double)node_total * (double)pts[cs].node_hot);
            cap_feat[cs] = (float)(nf / kg);
        }
    }
    int64_t best = 0;
    float best_saved = 0.f;                                                       // max_element: first maximum, entry 0 is 0 (:755-762)
    for (int64_t cs = 1; cs < steps; cs++) {
        const float tot = saved_topo[cs] + saved_feat[steps - 1 - cs];
        if (tot > best_saved) { best_saved = tot; best = cs; }
    }
    *node_capacity = (int32_t)(cap_feat[steps - 1 - best] + 1);
    *edge_capacity = (int32_t)(cap_topo[best] + 1);
    return LGN_OK;
}

// B200 placement model (no reference counterpart).  The reference decides "topology vs features" against
// PCIe transactions; on a B200 clique the decision that matters is how many of the hottest rows to REPLICATE on every
// GPU (served from local HBM) before the rest is partitioned over the clique (1/kg local, the rest over NVLink) and
// what stays on the host.  For n_repl replicated rows the per-GPU budget leaves room for (budget/row - n_repl) * kg
// partitioned rows; the expected time of one epoch's gathers is
//     t = H(n_repl)/bw_local + [H(n_repl + n_part) - H(n_repl)] * (1/kg/bw_local + (kg-1)/kg/bw_peer) + [H(N) - H(n_repl + n_part)]/bw_host
// with H(k) = presampled hotness of the k hottest rows + prior * k (a pseudo-count per row).  101 candidate splits are evaluated, the cheapest wins.
extern "C" int lgn_plan_hybrid(const uint32_t* af_sorted, int64_t n, int32_t dim, int64_t budget_bytes, int32_t kg,
                               double bw_local, double bw_peer, double bw_host, double prior, int64_t* n_repl_out,
                               int64_t* cap_out, double* est_cost, void* stream)
{
    if (!af_sorted || n <= 0 || dim <= 0 || budget_bytes <= 0 || kg <= 0 || kg > LGN_MAX_PARTS || !n_repl_out || !cap_out) return LGN_E_ARG;
    if (!(bw_local > 0) || !(bw_peer > 0) || !(bw_host > 0) || !(prior >= 0)) return LGN_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t row = (int64_t)dim * 4;
    const int64_t budget_rows = budget_bytes / row;
    const int n_pts = 101;
    DevBuf w, prefix, tmp;
    size_t tmp_bytes = 0;
    CK(w.alloc(n * 8)); CK(prefix.alloc(n * 8));
    auto* W = w.as<unsigned long long>();
    auto* H = prefix.as<unsigned long long>();
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, W, H, (int)n, s));
    CK(tmp.alloc(tmp_bytes));
    k_widen<<<1024, 256, 0, s>>>(af_sorted, W, n);
    CK(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, W, H, (int)n, s));
    auto h_at = [&](int64_t k, unsigned long long* out) -> int {   // H(k) = hotness of the k hottest rows
        *out = 0;
        if (k <= 0) return LGN_OK;
        if (k > n) k = n;
        CK(cudaMemcpyAsync(out, H + (k - 1), 8, cudaMemcpyDeviceToHost, s));
        return LGN_OK;
    };
    std::vector<int64_t> repl(n_pts), part(n_pts);
    std::vector<unsigned long long> h_repl(n_pts), h_all(n_pts);
    unsigned long long total = 0;
    const int64_t max_repl = budget_rows < n ? budget_rows : n;
    for (int i = 0; i < n_pts; i++) {
        repl[i] = kg == 1 ? 0 : max_repl * i / (n_pts - 1);
        int64_t room = (budget_rows - repl[i]) * kg;
        if (room > n - repl[i]) room = n - repl[i];
        part[i] = room < 0 ? 0 : room;
        int rc;
        if ((rc = h_at(repl[i], &h_repl[i]))) return rc;
        if ((rc = h_at(repl[i] + part[i], &h_all[i]))) return rc;
    }
    { int rc; if ((rc = h_at(n, &total))) return rc; }
    CK(cudaStreamSynchronize(s));
    const double c_mix = 1.0 / kg / bw_local + (double)(kg - 1) / kg / bw_peer;
    int best = 0;
    double best_t = -1.0;
    for (int i = 0; i < n_pts; i++) {
        // H(k) + prior * k: a row presampling never saw still gets `prior` expected reads per epoch, so rows are not pushed
        // to the host tier (12x the cost of a peer read) merely because ONE epoch of samples missed them
        const double hr = (double)h_repl[i] + prior * (double)repl[i];
        const double ha = (double)h_all[i] + prior * (double)(repl[i] + part[i]);
        const double ht = (double)total + prior * (double)n;
        const double t = hr / bw_local + (ha - hr) * c_mix + (ht - ha) / bw_host;
        if (best_t < 0 || t < best_t) { best_t = t; best = i; }
        if (kg == 1) break;
    }
    *n_repl_out = repl[best];
    *cap_out = repl[best] + (part[best] + kg - 1) / kg;
    if (*cap_out < 1) *cap_out = 1;
    if (est_cost) *est_cost = best_t * (double)row;      // expected seconds per presampled epoch if the bandwidths are bytes/s
    return LGN_OK;
}
