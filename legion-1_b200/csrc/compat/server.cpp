// server.cpp -- the reference's sampling server re-hosted on the B200 C-ABI.
//
// Mirrors, with its own code, GPUServer / GPURunner (reference Server.cu:43-364), the five operators
// (Operator.cu:10-123), GPUGraphStore's loader (GPUGraphStore.cu:145-469: ./meta_config, dataset files,
// `tid % P` seed partitioning, all-pairs P2P) and the cache build of GPUCache::CandidateSelection /
// CostModel / FillUp (GPUCache.cu:578-826).  Everything that touches the GPU goes through
// include/legion_b200.h; this file only owns host orchestration (threads, files, handshakes).
#include "Operator.h"
#include "Server.h"

#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../../include/legion_b200.h"

#define LGN_DIE(rc, what)                                                                                       \
    do {                                                                                                        \
        int rc_ = (rc);                                                                                         \
        if (rc_ != 0) {                                                                                         \
            fprintf(stderr, "legion_b200: %s failed: %s %s\n", what, lgn_error_string(rc_), lgn_last_cuda_error()); \
            exit(EXIT_FAILURE);   /* the reference exits on any CUDA error too (Kernels.cuh:14-22) */           \
        }                                                                                                       \
    } while (0)

#define CUDA_DIE(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) { fprintf(stderr, "legion_b200: %s: %s\n", #call, cudaGetErrorString(e_)); exit(EXIT_FAILURE); } \
    } while (0)

namespace {

// ---- per-GPU state handed to the operators through OpParams::memorypool (reference: GPUMemoryPool) ----
struct RunnerState {
    lgn_ctx* ctx = nullptr;
    int device = 0;
    int mode = 0;         // GPUMemoryPool::GetCurrentMode
    int iter = 0;         // GPUMemoryPool::GetIter
    int pipe = 0;         // GPUMemoryPool current_pipe_
    int batch_size = 0;   // IPCEnv::GetCurrentBatchsize(dev, mode)
};

// ---- IPCEnv (reference CUDA_IPC_Service.h:6-35): step arithmetic + the wire format ----
struct IPCEnv {
    lgn_ipc_server* ipc = nullptr;
    lgn_steps steps{};
    int epochs = 1, raw_batch = 0, parts = 0;
    int GetTrainStep() const { return steps.train_step; }
    int GetMaxStep() const { return steps.max_step; }
    int GetCurrentMode(int g) const { return lgn_mode_of_step(&steps, epochs, g); }
    int GetLocalBatchId(int g) const { return lgn_local_batch_id(&steps, epochs, g); }
    int GetCurrentBatchsize(int dev, int mode) const
    {
        return mode == LGN_MODE_TRAIN ? raw_batch : (mode == LGN_MODE_VALID ? steps.valid_batch[dev] : steps.test_batch[dev]);
    }
};

// ---- dataset in pinned + mapped host memory (GPUGraphStore::Load_Graph / Load_Feature) ----
struct Dataset {
    std::string path;
    int32_t batch = 0, n_nodes = 0, dim = 0, n_train = 0, n_valid = 0, n_test = 0, epochs = 1, partition_flag = 0;
    int64_t n_edges = 0, cache_memory = 0;
    int64_t *indptr_h = nullptr, *indptr_d = nullptr;
    int32_t *indices_h = nullptr, *indices_d = nullptr;
    float *feat_h = nullptr, *feat_d = nullptr;
    std::vector<int32_t> labels;
    std::vector<std::vector<int32_t>> ids[3], lab[3];   // [mode][partition]
};

void read_file(const std::string& file, void* dst, size_t bytes, bool required = true)
{
    int fd = open(file.c_str(), O_RDONLY);
    if (fd < 0) {
        if (required) { fprintf(stderr, "legion_b200: cannot open %s\n", file.c_str()); exit(EXIT_FAILURE); }
        return;
    }
    // parallel pread instead of the reference's single-threaded element-wise mmap copy loops
    const int nt = bytes > (64u << 20) ? 8 : 1;
    std::vector<std::thread> th;
    std::vector<size_t> got(nt, 0);
    const size_t chunk = (bytes + nt - 1) / nt;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=, &got]() {
            size_t off = t * chunk, end = off + chunk < bytes ? off + chunk : bytes;
            while (off < end) {
                ssize_t r = pread(fd, (char*)dst + off, end - off, (off_t)off);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) break;
                off += (size_t)r;
                got[t] += (size_t)r;
            }
        });
    for (auto& t : th) t.join();
    close(fd);
    size_t total = 0;
    for (size_t g : got) total += g;
    if (total != bytes) {   // a truncated file would otherwise run with uninitialised topology / features
        fprintf(stderr, "legion_b200: %s holds %zu bytes, meta_config implies %zu\n", file.c_str(), total, bytes);
        exit(EXIT_FAILURE);
    }
}

void load_dataset(Dataset& d, int parts)
{
    // ./meta_config: one line, 11 fields (GPUGraphStore::ReadMetaFIle, GPUGraphStore.cu:190-223)
    std::ifstream meta("./meta_config");
    if (!meta.is_open()) { fprintf(stderr, "legion_b200: unable to open ./meta_config\n"); exit(EXIT_FAILURE); }
    std::string line;
    std::getline(meta, line);
    std::istringstream iss(line);
    iss >> d.path >> d.batch >> d.n_nodes >> d.n_edges >> d.dim >> d.n_train >> d.n_valid >> d.n_test >> d.cache_memory >> d.epochs >> d.partition_flag;
    if (iss.fail() || d.batch <= 0 || d.n_nodes <= 0 || d.n_edges < 0 || d.dim <= 0 || d.n_train < 0 || d.n_valid < 0 || d.n_test < 0 || d.epochs < 0) {
        fprintf(stderr, "legion_b200: ./meta_config needs 11 fields: path batch nodes edges dim train valid test cache_bytes epochs partition\n");
        exit(EXIT_FAILURE);
    }
    std::cout << "Dataset path:       " << d.path << "\nRaw Batchsize:      " << d.batch << "\nGraph nodes num:    " << d.n_nodes
              << "\nGraph edges num:    " << d.n_edges << "\nFeature dim:        " << d.dim << "\nTraining set num:   " << d.n_train
              << "\nValidation set num: " << d.n_valid << "\nTesting set num:    " << d.n_test << "\nCache memory:       " << d.cache_memory
              << "\nTrain epoch:        " << d.epochs << "\nPartition?:         " << d.partition_flag << "\n";
    void *h = nullptr, *dv = nullptr;
    LGN_DIE(lgn_host_alloc_mapped(&h, &dv, (int64_t)(d.n_nodes + 1) * 8), "alloc indptr");
    d.indptr_h = (int64_t*)h; d.indptr_d = (int64_t*)dv;
    LGN_DIE(lgn_host_alloc_mapped(&h, &dv, d.n_edges * 4), "alloc indices");
    d.indices_h = (int32_t*)h; d.indices_d = (int32_t*)dv;
    LGN_DIE(lgn_host_alloc_mapped(&h, &dv, (int64_t)d.n_nodes * d.dim * 4), "alloc features");
    d.feat_h = (float*)h; d.feat_d = (float*)dv;
    read_file(d.path + "edge_src", d.indptr_h, (size_t)(d.n_nodes + 1) * 8);       // GPUGraphStore.cu:266-270
    read_file(d.path + "edge_dst", d.indices_h, (size_t)d.n_edges * 4);
    read_file(d.path + "features", d.feat_h, (size_t)d.n_nodes * d.dim * 4);
    d.labels.resize(d.n_nodes);
    read_file(d.path + "labels", d.labels.data(), (size_t)d.n_nodes * 4);
    std::vector<int32_t> raw[3] = {std::vector<int32_t>(d.n_train), std::vector<int32_t>(d.n_valid), std::vector<int32_t>(d.n_test)};
    const char* names[3] = {"trainingset", "validationset", "testingset"};
    for (int m = 0; m < 3; m++) read_file(d.path + names[m], raw[m].data(), raw[m].size() * 4);
    std::vector<int32_t> part_of;
    {   // optional xtrapulp partition of the training ids (GPUGraphStore.cu:300, 335-339)
        std::string pf = d.path + "partition_" + std::to_string(parts) + "_bn";
        struct stat st;
        if (d.partition_flag == 1 && stat(pf.c_str(), &st) == 0) { part_of.resize(d.n_nodes); read_file(pf, part_of.data(), (size_t)d.n_nodes * 4); }
    }
    for (int m = 0; m < 3; m++) {
        d.ids[m].assign(parts, {});
        d.lab[m].assign(parts, {});
        for (int32_t tid : raw[m]) {
            if (tid < 0 || tid >= d.n_nodes) { fprintf(stderr, "legion_b200: %s holds node id %d outside [0, %d)\n", names[m], tid, d.n_nodes); exit(EXIT_FAILURE); }
            int p = (m == 0 && !part_of.empty()) ? part_of[tid] : tid % parts;               // GPUGraphStore.cu:332-376
            if (p >= 0 && p < parts) { d.ids[m][p].push_back(tid); d.lab[m][p].push_back(d.labels[tid]); }
        }
    }
    std::cout << "Finish Reading All Files\n";
}

std::vector<int> env_fanout()
{
    std::vector<int> f;
    if (const char* e = getenv("LEGION_FANOUT")) {      // reference: hard-coded {25,10} (Server.cu:68-69)
        std::stringstream ss(e);
        std::string tok;
        while (std::getline(ss, tok, ',')) f.push_back(atoi(tok.c_str()));
    }
    if (f.empty()) f = {25, 10};
    return f;
}

// ================================================================= operators (Operator.cu:10-123)
class Batch_Generator : public Operator {
public:
    explicit Batch_Generator(int op_id) : op_id_(op_id) {}
    void run(OpParams* p) override
    {
        RunnerState* st = (RunnerState*)p->memorypool;
        LGN_DIE(lgn_batch_generate(st->ctx, p->stream, st->pipe, st->mode, st->batch_size, st->iter), "lgn_batch_generate");
        cudaEventRecord(p->event, p->stream);
    }
private:
    int op_id_;
};
class Random_Sampler : public Operator {
public:
    explicit Random_Sampler(int op_id) : op_id_(op_id) {}
    void run(OpParams* p) override
    {
        RunnerState* st = (RunnerState*)p->memorypool;
        LGN_DIE(lgn_select_pipe(st->ctx, st->pipe), "lgn_select_pipe");
        LGN_DIE(lgn_sample_hop(st->ctx, p->stream, op_id_ / 2 - 1, p->is_presc), "lgn_sample_hop");   // op ids 2,4,.. -> hop 0,1,..
        cudaEventRecord(p->event, p->stream);
    }
private:
    int op_id_;
};
class Feature_Extractor : public Operator {
public:
    explicit Feature_Extractor(int op_id) : op_id_(op_id) {}
    void run(OpParams* p) override
    {
        RunnerState* st = (RunnerState*)p->memorypool;
        LGN_DIE(lgn_select_pipe(st->ctx, st->pipe), "lgn_select_pipe");
        LGN_DIE(lgn_gather_segment(st->ctx, p->stream, (op_id_ - 1) / 2), "lgn_gather_segment");      // op ids 1,3,5 -> segment 0,1,2
    }
private:
    int op_id_;
};
class Cache_Planner : public Operator {
public:
    explicit Cache_Planner(int op_id) : op_id_(op_id) {}
    void run(OpParams* p) override
    {
        RunnerState* st = (RunnerState*)p->memorypool;
        LGN_DIE(lgn_select_pipe(st->ctx, st->pipe), "lgn_select_pipe");
        LGN_DIE(lgn_finish_batch(st->ctx, p->stream, p->is_presc), "lgn_finish_batch");
        cudaEventRecord(p->event, p->stream);
    }
private:
    int op_id_;
};
class Cache_Updater : public Operator {   // update_cache is an empty body in the reference too (Kernels.cu:786-805)
public:
    explicit Cache_Updater(int op_id) : op_id_(op_id) {}
    void run(OpParams* p) override { cudaEventRecord(p->event, p->stream); }
private:
    int op_id_;
};

}  // namespace

Operator* NewBatchGenerator(int op_id) { return new Batch_Generator(op_id); }
Operator* NewRandomSampler(int op_id) { return new Random_Sampler(op_id); }
Operator* NewFeatureExtractor(int op_id) { return new Feature_Extractor(op_id); }
Operator* NewCachePlanner(int op_id) { return new Cache_Planner(op_id); }
Operator* NewCacheUpdater(int op_id) { return new Cache_Updater(op_id); }

// ================================================================= GPURunner (Server.cu:167-364)
class GPURunner : public Runner {
public:
    void Initialize(RunnerParams* params) override
    {
        dev_ = params->device_id;
        CUDA_DIE(cudaSetDevice(dev_));
        env_ = (IPCEnv*)params->env;
        st_.ctx = (lgn_ctx*)params->cache;      // the per-GPU context built by the server
        st_.device = dev_;
        int lo = 0, hi = 0;
        CUDA_DIE(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_DIE(cudaStreamCreateWithPriority(&streams_[0], cudaStreamNonBlocking, hi));
        CUDA_DIE(cudaStreamCreateWithPriority(&streams_[1], cudaStreamNonBlocking, hi));
        const int hops = (int)params->fanout.size();
        op_num_ = (hops + 1) * 2 + 2;                                     // Server.cu:198-207
        ops_.resize(op_num_);
        ops_[0] = NewBatchGenerator(0);
        ops_[1] = NewFeatureExtractor(1);
        for (int i = 0; i < hops; i++) { ops_[2 * i + 2] = NewRandomSampler(2 * i + 2); ops_[2 * i + 3] = NewFeatureExtractor(2 * i + 3); }
        ops_[op_num_ - 2] = NewCachePlanner(op_num_ - 2);
        ops_[op_num_ - 1] = NewCacheUpdater(op_num_ - 1);
        params_.resize(op_num_);
        events_.resize(op_num_);
        for (int i = 0; i < op_num_; i++) {
            CUDA_DIE(cudaEventCreateWithFlags(&events_[i], cudaEventDisableTiming));
            params_[i] = OpParams{dev_, streams_[i % 2], events_[i], &st_, nullptr, nullptr, nullptr, env_, 0, false, params->in_memory};
        }
        for (int i = 0; i < hops; i++) params_[2 * i + 2].neighbor_count = params->fanout[i];
        use_ops_ = getenv("LEGION_RUNNER") && !strcmp(getenv("LEGION_RUNNER"), "ops");
    }
    void InitializeFeaturesBuffer(RunnerParams*) override {}   // the context owns worst-case feature buffers from the start

    void RunPreSc(RunnerParams* params) override   // Server.cu:284-299: ops 0,2,4,6 only
    {
        cudaSetDevice(dev_);
        st_.mode = LGN_MODE_TRAIN;
        st_.iter = params->global_batch_id;
        st_.batch_size = env_->GetCurrentBatchsize(dev_, LGN_MODE_TRAIN);
        st_.pipe = params->global_batch_id % LGN_PIPELINE_DEPTH;
        LGN_DIE(lgn_batch_generate(st_.ctx, streams_[st_.pipe], st_.pipe, st_.mode, st_.batch_size, st_.iter), "lgn_batch_generate");
        LGN_DIE(lgn_run_batch(st_.ctx, streams_[st_.pipe], 0, 1), "lgn_run_batch(presc)");
    }

    // Server.cu:301-328.  Batch i is enqueued BEFORE batch i-1 is waited for and posted, so the two slots
    // really overlap; the reference busy-polls batch i to completion before touching the next one.
    void RunOnce(RunnerParams* params) override
    {
        cudaSetDevice(dev_);
        const int g = params->global_batch_id;
        st_.mode = env_->GetCurrentMode(g);
        st_.iter = env_->GetLocalBatchId(g);
        st_.batch_size = env_->GetCurrentBatchsize(dev_, st_.mode);
        st_.pipe = pipe_;
        {   // Philox position: the epoch of the global batch id and a per-mode offset, so that no two mini-batches of a
            // run share a stream (the reference's minstd stream redraws the same neighbourhoods every epoch)
            const int per_epoch = env_->steps.train_step + env_->steps.valid_step;
            const bool test = g >= per_epoch * env_->epochs;
            const uint32_t epoch = test ? (uint32_t)env_->epochs : (uint32_t)(per_epoch > 0 ? g / per_epoch : 0);
            const uint32_t off = st_.mode == LGN_MODE_TRAIN ? 0u : (st_.mode == LGN_MODE_VALID ? (uint32_t)env_->steps.train_step : (uint32_t)per_epoch);
            LGN_DIE(lgn_set_epoch(st_.ctx, epoch + 1u, off), "lgn_set_epoch");      // epoch 0 is the presampling pass
        }
        LGN_DIE(lgn_ipc_server_wait(env_->ipc, dev_, pipe_), "IPCWait");
        if (use_ops_) {
            for (int i = 0; i < op_num_; i++) {      // the reference's operator DAG, event-chained on two streams
                if (i % 2 == 1) cudaStreamWaitEvent(streams_[1], events_[i - 1], 0);
                params_[i].is_presc = false;
                ops_[i]->run(&params_[i]);
            }
            CUDA_DIE(cudaStreamSynchronize(streams_[1]));
            CUDA_DIE(cudaStreamSynchronize(streams_[0]));
            LGN_DIE(lgn_sync_pipe_status(st_.ctx, pipe_), "mini-batch (device status)");
            LGN_DIE(lgn_ipc_server_post(env_->ipc, dev_, pipe_), "IPCPost");
        } else {
            LGN_DIE(lgn_batch_generate(st_.ctx, streams_[pipe_], pipe_, st_.mode, st_.batch_size, st_.iter), "lgn_batch_generate");
            LGN_DIE(lgn_run_batch(st_.ctx, streams_[pipe_], 1, 0), "lgn_run_batch");
            if (inflight_ >= 0) Publish(inflight_);
            inflight_ = pipe_;
        }
        pipe_ = (pipe_ + 1) % LGN_PIPELINE_DEPTH;
    }
    void Drain() { if (inflight_ >= 0) Publish(inflight_); inflight_ = -1; }
    void SyncAll() { cudaSetDevice(dev_); cudaStreamSynchronize(streams_[0]); cudaStreamSynchronize(streams_[1]); lgn_sync_pipe(st_.ctx, 0); lgn_sync_pipe(st_.ctx, 1); }

    void Finalize(RunnerParams*) override
    {
        Drain();
        LGN_DIE(lgn_ipc_server_wait(env_->ipc, dev_, (pipe_ + 1) % LGN_PIPELINE_DEPTH), "IPCWait(final)");   // Server.cu:330-334
        cudaSetDevice(dev_);
        for (Operator* op : ops_) delete op;
        ops_.clear();
        for (cudaEvent_t e : events_) cudaEventDestroy(e);
        events_.clear();
        cudaStreamDestroy(streams_[0]);
        cudaStreamDestroy(streams_[1]);
    }
    RunnerState st_;

private:
    void Publish(int pipe)
    {
        // a truncated block must never reach a trainer: capacity overflow (dedup table, id / edge / feature buffers) is fatal
        LGN_DIE(lgn_sync_pipe_status(st_.ctx, pipe), "mini-batch (device status)");
        LGN_DIE(lgn_ipc_server_post(env_->ipc, dev_, pipe), "IPCPost");
    }
    int dev_ = 0, op_num_ = 0, pipe_ = 0, inflight_ = -1;
    bool use_ops_ = false;
    IPCEnv* env_ = nullptr;
    cudaStream_t streams_[2];
    std::vector<Operator*> ops_;
    std::vector<OpParams> params_;
    std::vector<cudaEvent_t> events_;
};
Runner* NewGPURunner() { return new GPURunner(); }

// ================================================================= GPUServer (Server.cu:43-161)
class GPUServer : public Server {
public:
    void Initialize(int n) override
    {
        n_ = n;
        std::cout << "CUDA Device Count: " << n_ << "\n";
        LGN_DIE(lgn_enable_peer_access(n_), "EnableP2PAccess");                      // GPUGraphStore.cu:145-168
        load_dataset(ds_, n_);
        env_.parts = n_; env_.epochs = ds_.epochs; env_.raw_batch = ds_.batch;
        std::vector<int32_t> nt(n_), nv(n_), ns(n_);
        for (int i = 0; i < n_; i++) { nt[i] = (int32_t)ds_.ids[0][i].size(); nv[i] = (int32_t)ds_.ids[1][i].size(); ns[i] = (int32_t)ds_.ids[2][i].size(); }
        LGN_DIE(lgn_coordinate(nt.data(), nv.data(), ns.data(), n_, ds_.batch, ds_.epochs, &env_.steps), "Coordinate");
        std::cout << "Train Steps: " << env_.steps.train_step << "\nValid Steps: " << env_.steps.valid_step << "\nTest Steps: " << env_.steps.test_step << "\n";
        const int32_t steps[3] = {env_.steps.train_step, env_.steps.valid_step, env_.steps.test_step};
        LGN_DIE(lgn_ipc_server_create(n_, steps, &env_.ipc), "NewIPCEnv");
        fanout_ = env_fanout();
        const char* rng = getenv("LEGION_RNG");
        const char* seed = getenv("LEGION_SEED");
        ctx_.resize(n_);
        runners_.resize(n_);
        params_.resize(n_);
        int max_valid_test = 1;
        for (int i = 0; i < n_; i++) { if (env_.steps.valid_batch[i] > max_valid_test) max_valid_test = env_.steps.valid_batch[i]; if (env_.steps.test_batch[i] > max_valid_test) max_valid_test = env_.steps.test_batch[i]; }
        for (int i = 0; i < n_; i++) {
            lgn_config cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.device = i; cfg.part = 0; cfg.n_nodes = ds_.n_nodes; cfg.feat_dim = ds_.dim;
            cfg.batch_size = ds_.batch > max_valid_test ? ds_.batch : max_valid_test;
            cfg.n_hops = (int32_t)fanout_.size();
            for (size_t h = 0; h < fanout_.size(); h++) cfg.fanout[h] = fanout_[h];
            cfg.rng_mode = (rng && !strcmp(rng, "philox")) ? LGN_RNG_PHILOX : LGN_RNG_MINSTD;   // default: the reference's stream
            cfg.rng_seed = seed ? strtoull(seed, nullptr, 0) : 0;
            cfg.enable_hotness = 1;
            cfg.n_lanes = LGN_PIPELINE_DEPTH;
            LGN_DIE(lgn_create(&cfg, &ctx_[i]), "lgn_create");
            for (int m = 0; m < 3; m++) {                                              // GPU_Memory_Node_Storage.cu:52-94
                void *di = nullptr, *dl = nullptr;
                const size_t cnt = ds_.ids[m][i].size();
                LGN_DIE(lgn_device_alloc(&di, (int64_t)cnt * 4), "alloc ids");
                LGN_DIE(lgn_device_alloc(&dl, (int64_t)cnt * 4), "alloc labels");
                if (cnt) { LGN_DIE(lgn_copy_h2d(di, ds_.ids[m][i].data(), (int64_t)cnt * 4), "h2d ids"); LGN_DIE(lgn_copy_h2d(dl, ds_.lab[m][i].data(), (int64_t)cnt * 4), "h2d labels"); }
                LGN_DIE(lgn_bind_seeds(ctx_[i], m, (int32_t*)di, (int32_t*)dl, (int32_t)cnt), "lgn_bind_seeds");
                owned_.push_back({i, di}); owned_.push_back({i, dl});
            }
            LGN_DIE(lgn_bind_topology(ctx_[i], ds_.indptr_d, ds_.indices_d), "lgn_bind_topology");   // host CSR over UVA
            LGN_DIE(lgn_bind_features(ctx_[i], ds_.feat_d), "lgn_bind_features");
            LGN_DIE(lgn_ipc_server_publish(env_.ipc, i, ctx_[i], 1), "InitializeSamplesBuffer");
            params_[i] = new RunnerParams{i, fanout_, ctx_[i], nullptr, nullptr, &env_, 0, true};
            runners_[i] = new GPURunner();
            runners_[i]->Initialize(params_[i]);
        }
        std::cout << "Storage Initialized\n";
    }

    void PreSc(int cache_agg_mode) override
    {
        auto t0 = std::chrono::steady_clock::now();
        const int train_step = env_.steps.train_step;
        std::vector<std::thread> th;
        for (int i = 0; i < n_; i++)
            th.emplace_back([this, i, train_step]() {                                  // PreSCLoop, Server.cu:28-34
                for (int b = 0; b < train_step; b++) { params_[i]->global_batch_id = b; runners_[i]->RunPreSc(params_[i]); }
                runners_[i]->SyncAll();
            });
        for (auto& t : th) t.join();
        // hash dedup (large graphs): size the per-batch table for what presampling saw instead of the worst case,
        // so that it stays L2-resident (2.5x the largest batch inside the call; no-op for the direct map)
        // Presampling only saw TRAIN batches of raw_batch seeds; a valid/test batch may carry more seeds, so the estimate
        // is scaled by the largest batch any mode can have (overflow would be caught by Publish, but must not happen).
        for (int i = 0; i < n_; i++) {
            const int64_t seen = lgn_max_ids(ctx_[i], nullptr);
            int64_t largest = ds_.batch;
            if (env_.steps.valid_batch[i] > largest) largest = env_.steps.valid_batch[i];
            if (env_.steps.test_batch[i] > largest) largest = env_.steps.test_batch[i];
            const int64_t expect = (seen * largest + ds_.batch - 1) / ds_.batch;
            if (expect > 0) LGN_DIE(lgn_set_dedup_capacity(ctx_[i], expect), "lgn_set_dedup_capacity");
        }
        int kg = cache_agg_mode == 1 ? 2 : cache_agg_mode == 2 ? 4 : cache_agg_mode == 3 ? 8 : 1;   // GPUCache.cu:593-607
        if (kg > n_) {   // the reference computes Kc = 0 here and serves without any cache
            std::cout << "cache aggregate mode " << cache_agg_mode << " asks for " << kg << " GPUs per clique, only " << n_ << " present: using " << n_ << "\n";
            kg = n_;
        }
        const int kc = n_ / kg;
        std::cout << "NVLink Clique: " << kc << " GPU Per Clique: " << kg << std::endl;
        const int64_t N = ds_.n_nodes;
        for (int c = 0; c < kc; c++) {
            const int lead = c * kg;
            LGN_DIE(lgn_set_device(lead), "set device");
            uint32_t *nh = nullptr, *th2 = nullptr;
            LGN_DIE(lgn_hotness(ctx_[lead], &nh, &th2), "lgn_hotness");
            std::vector<int32_t> max_ids(kg);
            uint64_t trans = 0;
            // CandidateSelection: sum the clique's histograms.  NCCL all-reduce over NVLink (every GPU ends up with the
            // sum); the reference's leader-reads-peers loop (aggregate_access) remains as the fallback without NCCL.
            bool summed = false;
            if (kg > 1 && lgn_comm_available() && !getenv("LEGION_NO_NCCL")) {
                std::vector<int32_t> devs(kg);
                std::vector<uint32_t*> nptr(kg), tptr2(kg);
                for (int j = 0; j < kg; j++) { devs[j] = lead + j; LGN_DIE(lgn_hotness(ctx_[lead + j], &nptr[j], &tptr2[j]), "lgn_hotness"); }
                LGN_DIE(lgn_allreduce_u32_devices(kg, devs.data(), nptr.data(), N), "ncclAllReduce(node hotness)");
                LGN_DIE(lgn_allreduce_u32_devices(kg, devs.data(), tptr2.data(), N), "ncclAllReduce(topology hotness)");
                std::cout << "Hotness all-reduce: NCCL over " << kg << " GPUs\n";
                summed = true;
            }
            for (int j = 0; j < kg; j++) {
                if (j > 0 && !summed) {
                    uint32_t *pn = nullptr, *pt = nullptr;
                    LGN_DIE(lgn_hotness(ctx_[lead + j], &pn, &pt), "lgn_hotness");
                    LGN_DIE(lgn_set_device(lead), "set device");
                    LGN_DIE(lgn_accumulate_u32(nh, pn, N, nullptr), "aggregate_access");
                    LGN_DIE(lgn_accumulate_u32(th2, pt, N, nullptr), "aggregate_access");
                }
                max_ids[j] = lgn_max_ids(ctx_[lead + j], nullptr);
                int64_t tot[2];
                LGN_DIE(lgn_sampling_totals(ctx_[lead + j], nullptr, tot, 1), "lgn_sampling_totals");
                trans += (uint64_t)(tot[0] + tot[1]);   // one UVA read transaction per indptr pair and per neighbour id
            }
            LGN_DIE(lgn_set_device(lead), "set device");
            LGN_DIE(lgn_device_synchronize(), "sync");
            void *qf = nullptr, *qt = nullptr, *af = nullptr, *at = nullptr;
            LGN_DIE(lgn_device_alloc(&qf, N * 4), "alloc"); LGN_DIE(lgn_device_alloc(&qt, N * 4), "alloc");
            LGN_DIE(lgn_device_alloc(&af, N * 4), "alloc"); LGN_DIE(lgn_device_alloc(&at, N * 4), "alloc");
            LGN_DIE(lgn_hot_order(nh, N, (int32_t*)qf, (uint32_t*)af, nullptr), "hot order (features)");
            LGN_DIE(lgn_hot_order(th2, N, (int32_t*)qt, (uint32_t*)at, nullptr), "hot order (topology)");
            int32_t ncap = 0, ecap = 0;
            int64_t n_repl = 0;
            const int64_t feat_bytes = N * ds_.dim * 4, topo_bytes = 8 * N + 4 * ds_.n_edges;
            const char* pl = getenv("LEGION_PLACEMENT");
            const bool hybrid = !(pl && !strcmp(pl, "reference"));
            bool topo_replicated = false;
            if (hybrid) {
                // B200 placement.  Topology: a full copy in every GPU's HBM when it takes at most a quarter of the
                // budget (7.4 GB for papers100M), otherwise the reference's partition with the capacity its cost model picks.
                // Features: lgn_plan_hybrid splits the remaining budget into rows replicated on every GPU of the clique, rows
                // partitioned over it and rows left on the host, by expected gather time over the three tiers.
                int64_t feat_budget = ds_.cache_memory;
                if (topo_bytes + 8 <= ds_.cache_memory / 4) {
                    topo_replicated = true;
                    feat_budget -= topo_bytes + 8;
                } else {
                    int32_t ncap_ref = 0;
                    LGN_DIE(lgn_cost_model((uint32_t*)af, (uint32_t*)at, (int32_t*)qt, ds_.indptr_d, N, ds_.dim, ds_.cache_memory, kg, trans,
                                           max_ids.data(), train_step, &ncap_ref, &ecap, nullptr), "CostModel(topology)");
                    const int64_t avg_adj = 8 + 4 * (ds_.n_edges / (N > 0 ? N : 1) + 1);
                    const int64_t topo_share = (int64_t)ecap * avg_adj;
                    if (topo_share > ds_.cache_memory / 2) ecap = (int32_t)(ds_.cache_memory / 2 / avg_adj);
                    feat_budget -= (int64_t)ecap * avg_adj;
                }
                if (feat_budget < ds_.dim * 4) feat_budget = ds_.dim * 4;
                int64_t cap64 = 0;
                double cost = 0;
                const double bw_local = 3272.0, bw_peer = 640.0, bw_host = 50.0;     // GB/s payload per GPU: HBM copy / 2, measured NVLink gather, PCIe Gen5 zero-copy
                LGN_DIE(lgn_plan_hybrid((uint32_t*)af, N, ds_.dim, feat_budget, kg, bw_local, bw_peer, bw_host, /*prior=*/0.5, &n_repl, &cap64, &cost, nullptr), "lgn_plan_hybrid");
                ncap = (int32_t)cap64;
                if (topo_replicated) ecap = 0;
                std::cout << "Placement: hybrid, " << n_repl << " hottest rows replicated, " << (cap64 - n_repl) * kg << " partitioned over " << kg
                          << " GPU(s), topology " << (topo_replicated ? "replicated in HBM" : "partitioned") << "\n";
            } else if (ds_.cache_memory * kg > feat_bytes + topo_bytes + (int64_t)kg * (8 + ds_.dim * 4)) {
                // Everything fits.  The reference's CostModel leaves both capacity tables at 0 in this case
                // (GPUCache.cu:744-751 only fill them while a tier does NOT fit) and would cache one row; with
                // 180 GB per B200 this is the normal case, so it is handled explicitly: cache all of both tiers.
                ncap = ecap = (int32_t)((N + kg - 1) / kg);
                std::cout << "Cost model: whole graph fits, caching everything\n";
            } else {
                LGN_DIE(lgn_cost_model((uint32_t*)af, (uint32_t*)at, (int32_t*)qt, ds_.indptr_d, N, ds_.dim, ds_.cache_memory, kg, trans,
                                       max_ids.data(), train_step, &ncap, &ecap, nullptr), "CostModel");
            }
            std::cout << "Feat capacity " << ncap << " topo capacity " << ecap << std::endl;
            // FillUp (GPUCache.cu:769-826): shard j of the clique lives on GPU lead+j
            std::vector<const float*> fshard(kg);
            std::vector<const int64_t*> tptr(kg);
            std::vector<const int32_t*> tidx(kg);
            std::vector<int32_t*> fslot(kg), tslot(kg);
            std::vector<void*> fcmap(kg, nullptr);
            for (int j = 0; j < kg; j++) {
                const int dev = lead + j;
                LGN_DIE(lgn_set_device(dev), "set device");
                void *oq = qf, *ot = qt;
                if (j > 0) {   // replicate the two orders next to the shard builder
                    LGN_DIE(lgn_device_alloc(&oq, N * 4), "alloc"); LGN_DIE(lgn_device_alloc(&ot, N * 4), "alloc");
                    LGN_DIE(lgn_copy_d2d(oq, qf, N * 4), "copy order"); LGN_DIE(lgn_copy_d2d(ot, qt, N * 4), "copy order");
                }
                void *shard = nullptr, *fs = nullptr, *ts = nullptr, *tip = nullptr, *tix = nullptr, *cm = nullptr;
                LGN_DIE(lgn_device_alloc(&shard, (int64_t)ncap * ds_.dim * 4), "alloc feature shard");
                if (!hybrid) LGN_DIE(lgn_device_alloc(&fs, N * 4), "alloc");
                LGN_DIE(lgn_device_alloc(&ts, N * 4), "alloc");
                if (hybrid) {
                    // compact placement map (37 MB for papers100M: L2-resident) instead of int32 slot_of[N]; rows of a class in
                    // node-id order, so FeatFillUp is one streaming pass over the feature matrix
                    int64_t n_part = ((int64_t)ncap - n_repl) * kg;
                    if (n_part > N - n_repl) n_part = N - n_repl;
                    LGN_DIE(lgn_device_alloc(&cm, lgn_cmap_bytes(N)), "alloc placement map");
                    LGN_DIE(lgn_place_compact((int32_t*)oq, N, n_repl, n_part, cm, nullptr), "InitPair (compact)");
                    LGN_DIE(lgn_fill_feature_shard_compact(cm, N, n_repl, kg, j, ds_.feat_d, ds_.dim, (float*)shard, ncap, nullptr), "FeatFillUp");
                } else {
                    LGN_DIE(lgn_fill_feature_shard_hybrid((int32_t*)oq, N, ncap, kg, j, n_repl, ds_.feat_d, ds_.dim, (float*)shard, nullptr), "FeatFillUp");
                    LGN_DIE(lgn_place_hybrid((int32_t*)oq, N, ncap, kg, n_repl, j, (int32_t*)fs, nullptr), "InitPair");
                }
                if (topo_replicated) {      // full CSR copy in this GPU's HBM: the sampler never leaves the device
                    LGN_DIE(lgn_device_alloc(&tip, (int64_t)(N + 1) * 8), "alloc");
                    LGN_DIE(lgn_device_alloc(&tix, (ds_.n_edges > 0 ? ds_.n_edges : 1) * 4), "alloc");
                    LGN_DIE(lgn_copy_h2d(tip, ds_.indptr_h, (int64_t)(N + 1) * 8), "copy indptr");
                    LGN_DIE(lgn_copy_h2d(tix, ds_.indices_h, ds_.n_edges * 4), "copy indices");
                } else {
                    LGN_DIE(lgn_place((int32_t*)ot, N, ecap, kg, (int32_t*)ts, nullptr), "InitIndexPair");
                    LGN_DIE(lgn_device_alloc(&tip, (int64_t)(ecap + 1) * 8), "alloc");
                    int64_t cnt = 0;
                    LGN_DIE(lgn_fill_topo_shard((int32_t*)ot, N, ecap, kg, j, ds_.indptr_d, ds_.indices_d, (int64_t*)tip, nullptr, &cnt, nullptr), "GetNeighborCount");
                    LGN_DIE(lgn_device_alloc(&tix, (cnt > 0 ? cnt : 1) * 4), "alloc");
                    LGN_DIE(lgn_fill_topo_shard((int32_t*)ot, N, ecap, kg, j, ds_.indptr_d, ds_.indices_d, (int64_t*)tip, (int32_t*)tix, &cnt, nullptr), "TopoFillUp");
                }
                LGN_DIE(lgn_device_synchronize(), "sync");
                fshard[j] = (float*)shard; fslot[j] = (int32_t*)fs; tslot[j] = (int32_t*)ts; tptr[j] = (int64_t*)tip; tidx[j] = (int32_t*)tix;
                fcmap[j] = cm;
                for (void* q : {shard, fs, ts, tip, tix, cm}) if (q) owned_.push_back({dev, q});
                if (j > 0) { lgn_device_free(oq); lgn_device_free(ot); }
            }
            for (int j = 0; j < kg; j++) {   // every GPU of the clique sees all shards (P2P) and its own replica of the maps
                LGN_DIE(lgn_set_part(ctx_[lead + j], j), "lgn_set_part");
                const bool resident = hybrid && (n_repl >= N || (kg == 1 && (int64_t)ncap >= N));   // whole matrix in this GPU's shard, id order
                if (resident) LGN_DIE(lgn_bind_feature_cache_compact(ctx_[lead + j], kg, fshard.data(), nullptr, N, ncap), "bind resident features");
                else if (hybrid) LGN_DIE(lgn_bind_feature_cache_compact(ctx_[lead + j], kg, fshard.data(), fcmap[j], n_repl, ncap), "bind feature cache (compact)");
                else LGN_DIE(lgn_bind_feature_cache(ctx_[lead + j], kg, fshard.data(), fslot[j], ncap), "bind feature cache");
                if (topo_replicated) LGN_DIE(lgn_bind_topology(ctx_[lead + j], tptr[j], tidx[j]), "bind topology (HBM copy)");
                else LGN_DIE(lgn_bind_topology_cache(ctx_[lead + j], kg, tptr.data(), tidx.data(), tslot[j], ecap), "bind topology cache");
            }
            LGN_DIE(lgn_set_device(lead), "set device");
            lgn_device_free(qf); lgn_device_free(qt); lgn_device_free(af); lgn_device_free(at);
        }
        std::cout << "Finish load feature cache\nFinish load topology cache\n";
        double t = std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "First epoch cost: " << t << " s\n";
        std::cout << "System is ready for serving" << std::endl;
    }

    void Run() override
    {
        const int max_step = env_.steps.max_step;
        std::vector<std::thread> th;
        for (int i = 0; i < n_; i++)
            th.emplace_back([this, i, max_step]() {                                    // RunnerLoop, Server.cu:36-41
                for (int g = 0; g < max_step; g++) { params_[i]->global_batch_id = g; runners_[i]->RunOnce(params_[i]); }
                runners_[i]->Drain();
            });
        for (auto& t : th) t.join();
    }

    void Finalize() override
    {
        for (int i = 0; i < n_; i++) runners_[i]->Finalize(params_[i]);
        for (int i = 0; i < n_; i++) lgn_destroy(ctx_[i]);
        for (auto& a : owned_) { lgn_set_device(a.first); lgn_device_free(a.second); }   // seed sets, cache shards, slot tables
        owned_.clear();
        for (int i = 0; i < n_; i++) { delete runners_[i]; delete params_[i]; }
        lgn_ipc_server_destroy(env_.ipc);
        lgn_host_free(ds_.indptr_h); lgn_host_free(ds_.indices_h); lgn_host_free(ds_.feat_h);
        std::cout << "Server Stopped\n";
    }

private:
    int n_ = 0;
    Dataset ds_;
    IPCEnv env_;
    std::vector<int> fanout_;
    std::vector<lgn_ctx*> ctx_;
    std::vector<GPURunner*> runners_;
    std::vector<RunnerParams*> params_;
    std::vector<std::pair<int, void*>> owned_;   // (device, allocation) released in Finalize
};
Server* NewGPUServer() { return new GPUServer(); }
