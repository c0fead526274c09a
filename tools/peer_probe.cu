// peer_probe.cu -- measurement tool, not part of the product: what does a random row gather out of a PEER GPU's
// memory sustain over NVLink/NVSwitch on this box, as a function of table size (remote L2 vs remote DRAM), kernel
// variant (128-bit LDG warp-per-row vs cp.async.bulk thread-per-row) and grid size?  The number is the practical
// ceiling of the peer tier in the hit-mix roofline.   usage: peer_probe [n_gpus=2]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <unistd.h>
#include <sys/stat.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16; return x; }

// row r of the launch reads table row idx(r) of peer (r % n_peers): spreads the rows over all peers like rank-round-robin sharding
struct Tabs { const float* tab[8]; int n; };

template <int U>
__global__ void __launch_bounds__(256) k_ldg(const __grid_constant__ Tabs tb, long long rows_per_tab, int n_rows, int row_f, float* __restrict__ out, uint32_t salt)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    const int nvec = row_f >> 2;
    for (int r0 = warp * U; r0 < n_rows; r0 += n_warps * U) {
        uint4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = r0 + u;
            if (r < n_rows) {
                const uint32_t h = mix((uint32_t)r ^ salt);
                const uint4* s = reinterpret_cast<const uint4*>(tb.tab[r % tb.n] + (long long)(h % (uint32_t)rows_per_tab) * row_f);
#pragma unroll
                for (int k = 0; k < 2; k++) if (lane + 32 * k < nvec) v[u][k] = __ldg(s + lane + 32 * k);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = r0 + u;
            if (r < n_rows) {
                uint4* d = reinterpret_cast<uint4*>(out + (long long)r * row_f);
#pragma unroll
                for (int k = 0; k < 2; k++) if (lane + 32 * k < nvec) d[lane + 32 * k] = v[u][k];
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_bulk(const __grid_constant__ Tabs tb, long long rows_per_tab, int n_rows, int row_f, float* __restrict__ out, uint32_t salt)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int t = threadIdx.x;
    const uint32_t row_bytes = (uint32_t)row_f * 4u;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)blockDim.x * row_bytes);
    const uint32_t my_buf = (uint32_t)__cvta_generic_to_shared(smem + (size_t)t * row_bytes);
    const uint32_t my_bar = (uint32_t)__cvta_generic_to_shared(bars + t);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(my_bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    for (int r = blockIdx.x * blockDim.x + t; r < n_rows; r += gridDim.x * blockDim.x) {
        const uint32_t h = mix((uint32_t)r ^ salt);
        const float* src = tb.tab[r % tb.n] + (long long)(h % (uint32_t)rows_per_tab) * row_f;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(my_bar), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(my_buf), "l"(src), "r"(row_bytes), "r"(my_bar) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(my_bar), "r"(phase) : "memory");
        phase ^= 1u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (long long)r * row_f), "r"(my_buf), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// two rows in flight per thread (double-buffered staging): twice the bytes in flight per CTA at the same thread count
__global__ void __launch_bounds__(256) k_bulk2(const __grid_constant__ Tabs tb, long long rows_per_tab, int n_rows, int row_f, float* __restrict__ out, uint32_t salt)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int t = threadIdx.x;
    const uint32_t row_bytes = (uint32_t)row_f * 4u;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)blockDim.x * 2 * row_bytes);
    uint32_t buf[2], bar[2], phase[2] = {0, 0};
    for (int b = 0; b < 2; b++) {
        buf[b] = (uint32_t)__cvta_generic_to_shared(smem + ((size_t)t * 2 + b) * row_bytes);
        bar[b] = (uint32_t)__cvta_generic_to_shared(bars + t * 2 + b);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[b]));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int stride = gridDim.x * blockDim.x;
    auto issue = [&](int r, int b) {
        const uint32_t h = mix((uint32_t)r ^ salt);
        const float* src = tb.tab[r % tb.n] + (long long)(h % (uint32_t)rows_per_tab) * row_f;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[b]), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(buf[b]), "l"(src), "r"(row_bytes), "r"(bar[b]) : "memory");
    };
    int r = blockIdx.x * blockDim.x + t;
    if (r < n_rows) issue(r, 0);
    int b = 0;
    while (r < n_rows) {
        const int rn = r + stride;
        if (rn < n_rows) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // buffer b^1's previous store has drained
            issue(rn, b ^ 1);
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar[b]), "r"(phase[b]) : "memory");
        phase[b] ^= 1u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (long long)r * row_f), "r"(buf[b]), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        r = rn;
        b ^= 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// product-like variants of the LDG gather: HINT = cache-hinted loads/stores as in gather.cu (ld.global.nc.L1::no_allocate
// .L2::cache_hint evict_first), SKEW = half of the rows come from the first 64 Ki rows of each table and are the same
// on every reader (hot set), LOOKUP = a dependent random 4-byte read of a local array precedes every row
template <bool HINT, bool SKEW, bool LOOKUP>
__global__ void __launch_bounds__(256) k_like(const __grid_constant__ Tabs tb, long long rows_per_tab, int n_rows, int row_f, float* __restrict__ out,
                                              uint32_t salt, const uint32_t* __restrict__ lut, uint32_t lut_n)
{
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (gridDim.x * 256) >> 5;
    const int nvec = row_f >> 2;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int r0 = warp * 32; r0 < n_rows; r0 += n_warps * 32) {
        // lane l resolves row r0+l (like the product: one lookup per lane, then 4 rows at a time per warp)
        const int r = r0 + lane;
        const uint4* src = nullptr;
        if (r < n_rows) {
            uint32_t h = mix((uint32_t)r ^ salt);
            if (LOOKUP) h ^= __ldg(lut + (h % lut_n));
            long long row = (long long)(h % (uint32_t)rows_per_tab);
            if (SKEW && (h & 0x10000u)) row = (long long)(mix((uint32_t)r) & 0xffffu) % rows_per_tab;   // salt-free: same hot rows everywhere
            src = reinterpret_cast<const uint4*>(tb.tab[(h >> 20) % tb.n] + row * row_f);
        }
        const int rows = min(32, n_rows - r0);
        for (int rr = 0; rr < rows; rr += U) {
            uint4 v[U];
            const uint4* sp[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                sp[u] = reinterpret_cast<const uint4*>(__shfl_sync(0xffffffffu, (unsigned long long)src, (rr + u) & 31));
                if (rr + u >= rows) sp[u] = nullptr;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (sp[u] && lane < nvec) {
                    if (HINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                           : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(sp[u] + lane), "l"(pol));
                    else v[u] = __ldg(sp[u] + lane);
                }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (sp[u] && lane < nvec) {
                    uint4* d = reinterpret_cast<uint4*>(out + (long long)(r0 + rr + u) * row_f) + lane;
                    if (HINT) asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(d), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w), "l"(pol) : "memory");
                    else *d = v[u];
                }
        }
    }
}

__global__ void k_fill(float* p, long long n, float v) { for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v + (float)(i & 1023); }

// ---- multi-process mode: one process per GPU, peer tables mapped through CUDA IPC handles (what the torchrun
// deployment does); file-based rendezvous in a scratch directory.   usage: peer_probe ipc <rank> <world> <dir> [row floats]
static void file_barrier(const std::string& dir, int rank, int world, int& seq)
{
    const std::string me = dir + "/b" + std::to_string(seq) + "_" + std::to_string(rank);
    FILE* f = fopen(me.c_str(), "w"); if (f) fclose(f);
    for (int p = 0; p < world; p++) {
        const std::string other = dir + "/b" + std::to_string(seq) + "_" + std::to_string(p);
        struct stat sb; int spins = 0;
        while (stat(other.c_str(), &sb) != 0) { usleep(200); if (++spins > 300000) { fprintf(stderr, "barrier timeout\n"); exit(3); } }
    }
    seq++;
}

static int ipc_main(int argc, char** argv)
{
    const int rank = atoi(argv[2]), world = atoi(argv[3]);
    const std::string dir = argv[4];
    const int row_f = argc > 5 ? atoi(argv[5]) : 128;
    const int n_rows = 400000, reps = 6;
    // PROBE_ALLOC_MB: size of each GPU's table ALLOCATION (default 8 GiB); e.g. 2441 reproduces the product's 2.56 GB shards
    const size_t max_table = getenv("PROBE_ALLOC_MB") ? (size_t)atoll(getenv("PROBE_ALLOC_MB")) << 20 : (size_t)8 << 30;
    int seq = 0;
    CK(cudaSetDevice(rank));
    float *table, *out;
    if (getenv("PROBE_PREALLOC_MB")) {   // stands in for the dataset a product process holds before it allocates its shard
        void* dummy; CK(cudaMalloc(&dummy, (size_t)atoll(getenv("PROBE_PREALLOC_MB")) << 20));
    }
    const int n_kinds = getenv("PROBE_KINDS") ? atoi(getenv("PROBE_KINDS")) : 6;
    CK(cudaMalloc(&table, max_table));
    CK(cudaMalloc(&out, (size_t)n_rows * row_f * 4));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, rank));
    k_fill<<<pr.multiProcessorCount * 8, 256, 0, s>>>(table, (long long)(max_table / 4), (float)rank);
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaStreamSynchronize(s));
    const uint32_t lut_n = 40u << 20;                      // 160 MB of 4-byte entries: the product's slot table at 40 M nodes
    uint32_t* lut; CK(cudaMalloc(&lut, (size_t)lut_n * 4));
    k_fill<<<pr.multiProcessorCount * 8, 256, 0, s>>>((float*)lut, (long long)lut_n, 1.0f);
    CK(cudaStreamSynchronize(s));
    cudaIpcMemHandle_t mine; CK(cudaIpcGetMemHandle(&mine, table));
    { const std::string tmp = dir + "/h" + std::to_string(rank) + ".tmp", fin = dir + "/h" + std::to_string(rank);
      FILE* f = fopen(tmp.c_str(), "wb"); fwrite(&mine, sizeof(mine), 1, f); fclose(f); rename(tmp.c_str(), fin.c_str()); }
    file_barrier(dir, rank, world, seq);
    Tabs tb; tb.n = 0;
    std::vector<void*> opened;
    for (int p = 0; p < world; p++) {
        if (p == rank) continue;
        cudaIpcMemHandle_t h; const std::string fin = dir + "/h" + std::to_string(p);
        struct stat sb; while (stat(fin.c_str(), &sb) != 0) usleep(200);
        FILE* f = fopen(fin.c_str(), "rb"); if (fread(&h, sizeof(h), 1, f) != 1) { fprintf(stderr, "short handle\n"); return 3; } fclose(f);
        void* ptr; CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        opened.push_back(ptr); tb.tab[tb.n++] = (const float*)ptr;
    }
    if (rank == 0) printf("# peer_probe ipc: %d processes, rows of %d B, %d rows per launch; peer tables mapped with cudaIpcOpenMemHandle\n", world, row_f * 4, n_rows);
    const size_t sizes[] = {(size_t)96 << 20, (size_t)1 << 30, max_table};
    for (size_t sz : sizes) {
        if (sz > max_table) continue;
        const long long rows_per_tab = (long long)(sz / ((size_t)row_f * 4));
        for (int kind = 0; kind < n_kinds; kind++) {
            file_barrier(dir, rank, world, seq);
            for (int rep = -1; rep < reps; rep++) {
                if (rep == 0) CK(cudaEventRecord(e0, s));
                const uint32_t salt = 0x9e3779b9u * (uint32_t)(rep + 2) + (uint32_t)rank;
                const int g3 = pr.multiProcessorCount * 3;
                if (kind == 0) k_ldg<4><<<g3, 256, 0, s>>>(tb, rows_per_tab, n_rows, row_f, out, salt);
                else if (kind == 1) k_bulk<<<pr.multiProcessorCount * 2, 192, (size_t)192 * (row_f * 4 + 8), s>>>(tb, rows_per_tab, n_rows, row_f, out, salt);
                else if (kind == 2) k_like<true, false, false><<<g3, 256, 0, s>>>(tb, rows_per_tab, n_rows, row_f, out, salt, lut, lut_n);
                else if (kind == 3) k_like<false, true, false><<<g3, 256, 0, s>>>(tb, rows_per_tab, n_rows, row_f, out, salt, lut, lut_n);
                else if (kind == 4) k_like<false, false, true><<<g3, 256, 0, s>>>(tb, rows_per_tab, n_rows, row_f, out, salt, lut, lut_n);
                else k_like<true, true, true><<<g3, 256, 0, s>>>(tb, rows_per_tab, n_rows, row_f, out, salt, lut, lut_n);
            }
            CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("ipc rank %d  table %5zu MiB  %-18s %8.1f GB/s\n", rank, sz >> 20, kind == 0 ? "ldg U4 x3" : kind == 1 ? "bulk 1row x192 x2" : kind == 2 ? "like: hints" : kind == 3 ? "like: skew" : kind == 4 ? "like: lookup" : "like: all three", (double)n_rows * row_f * 4 * reps / (ms * 1e-3) / 1e9);
            fflush(stdout);
        }
    }
    file_barrier(dir, rank, world, seq);
    for (void* q : opened) cudaIpcCloseMemHandle(q);
    file_barrier(dir, rank, world, seq);
    return 0;
}

struct Dev { float* table; float* out; cudaStream_t s; cudaEvent_t e0, e1; int n_sm; };

int main(int argc, char** argv)
{
    if (argc >= 5 && std::string(argv[1]) == "ipc") return ipc_main(argc, argv);
    int G = argc > 1 ? atoi(argv[1]) : 2, have = 0;
    CK(cudaGetDeviceCount(&have));
    if (have < G) G = have;
    if (G < 1) { fprintf(stderr, "no GPU\n"); return 1; }
    const int row_f = argc > 2 ? atoi(argv[2]) : 128;           // floats per row (512 B)
    const int n_rows = 400000, reps = 6;
    const size_t max_table = (size_t)8 << 30;
    std::vector<Dev> d(G);
    for (int g = 0; g < G; g++) {
        CK(cudaSetDevice(g));
        for (int p = 0; p < G; p++) if (p != g) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError(); }
        CK(cudaMalloc(&d[g].table, max_table));
        CK(cudaMalloc(&d[g].out, (size_t)n_rows * row_f * 4));
        CK(cudaStreamCreateWithFlags(&d[g].s, cudaStreamNonBlocking));
        CK(cudaEventCreate(&d[g].e0)); CK(cudaEventCreate(&d[g].e1));
        cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, g)); d[g].n_sm = pr.multiProcessorCount;
        k_fill<<<d[g].n_sm * 8, 256, 0, d[g].s>>>(d[g].table, (long long)(max_table / 4), (float)g);
        CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_bulk2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    for (int g = 0; g < G; g++) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
    printf("# peer_probe: %d GPU(s), rows of %d B, %d rows per launch, %d launches per point; GB/s = payload read per reading GPU\n", G, row_f * 4, n_rows, reps);
    printf("%-8s %-10s %-22s %6s %10s\n", "readers", "table", "variant", "ctas", "GB/s/GPU");
    const size_t sizes[] = {(size_t)96 << 20, (size_t)1 << 30, (size_t)8 << 30};
    struct Var { const char* name; int kind; int ctas; int threads; };
    const Var vars[] = {{"ldg U4", 0, 3, 256}, {"ldg U4", 0, 8, 256}, {"ldg U8", 1, 3, 256}, {"ldg U8", 1, 8, 256},
                        {"bulk 1row x192thr", 2, 1, 192}, {"bulk 1row x192thr", 2, 2, 192}, {"bulk 2rows x160thr", 3, 1, 160}, {"bulk 2rows x96thr", 3, 2, 96}};
    // target: 0 = local table (HBM baseline), 1 = all peers
    for (int target = 0; target < 2; target++) {
        if (target == 1 && G < 2) break;
        for (int readers = 1; readers <= G; readers += (G > 1 ? G - 1 : 1)) {     // one reader, then all GPUs reading at once
            for (size_t sz : sizes) {
                const long long rows_per_tab = (long long)(sz / ((size_t)row_f * 4));
                for (const Var& v : vars) {
                    float worst_ms = 0;
                    for (int rep = -1; rep < reps; rep++) {
                        if (rep == 0) for (int g = 0; g < readers; g++) { CK(cudaSetDevice(g)); CK(cudaEventRecord(d[g].e0, d[g].s)); }
                        for (int g = 0; g < readers; g++) {
                            CK(cudaSetDevice(g));
                            Tabs tb; tb.n = 0;
                            if (target == 0) tb.tab[tb.n++] = d[g].table;
                            else for (int p = 0; p < G; p++) if (p != g) tb.tab[tb.n++] = d[p].table;
                            const uint32_t salt = 0x9e3779b9u * (uint32_t)(rep + 2) + (uint32_t)g;
                            const int grid = d[g].n_sm * v.ctas;
                            if (v.kind == 0) k_ldg<4><<<grid, 256, 0, d[g].s>>>(tb, rows_per_tab, n_rows, row_f, d[g].out, salt);
                            else if (v.kind == 1) k_ldg<8><<<grid, 256, 0, d[g].s>>>(tb, rows_per_tab, n_rows, row_f, d[g].out, salt);
                            else if (v.kind == 2) k_bulk<<<grid, v.threads, (size_t)v.threads * (row_f * 4 + 8), d[g].s>>>(tb, rows_per_tab, n_rows, row_f, d[g].out, salt);
                            else k_bulk2<<<grid, v.threads, (size_t)v.threads * 2 * (row_f * 4 + 8), d[g].s>>>(tb, rows_per_tab, n_rows, row_f, d[g].out, salt);
                        }
                    }
                    for (int g = 0; g < readers; g++) { CK(cudaSetDevice(g)); CK(cudaEventRecord(d[g].e1, d[g].s)); }
                    for (int g = 0; g < readers; g++) {
                        CK(cudaSetDevice(g)); CK(cudaEventSynchronize(d[g].e1));
                        float ms; CK(cudaEventElapsedTime(&ms, d[g].e0, d[g].e1));
                        if (ms > worst_ms) worst_ms = ms;
                    }
                    CK(cudaGetLastError());
                    const double gbps = (double)n_rows * row_f * 4 * reps / (worst_ms * 1e-3) / 1e9;
                    printf("%-8s %-10s %-22s %6d %10.1f\n", target == 0 ? "local" : (readers == 1 ? "1 peer-rd" : "all"), sz >= ((size_t)1 << 30) ? (sz == ((size_t)1 << 30) ? "1 GiB" : "8 GiB") : "96 MiB",
                           v.name, v.ctas, gbps);
                    fflush(stdout);
                }
            }
            if (target == 0) break;    // local baseline: one GPU is enough
        }
    }
    return 0;
}
