// legion_operator.h -- the operator plugin boundary of the reference, restated for the B200 library.
//
// The reference drives one GPU with a DAG of five operator kinds; every operator receives the same
// parameter block (reference Operator.h:4-16) and exposes one virtual entry point (Operator.h:18-21).
// Code written against that header keeps compiling against this one: the type names, the member names
// and their order, and the five factory functions (Operator.h:23-27) are the contract.  What the
// members MEAN on this side is documented per member below; the implementations live in server.cpp and
// are thin calls into the C-ABI of include/legion_b200.h.
#ifndef LEGION_B200_COMPAT_OPERATOR_H
#define LEGION_B200_COMPAT_OPERATOR_H

#include <cuda_runtime.h>

struct OpParams {
    // GPU this operator instance belongs to; one runner (host thread) per device.
    int          device_id;
    // Stream the operator enqueues on.  All five operators of one runner share it, as in the reference
    // (Server.cu:176-207); the library forks its own gather lanes off this stream internally.
    cudaStream_t stream;
    // Recorded by the runner after the operator returns; kept for source compatibility only.
    cudaEvent_t  event;
    // reference: GPUMemoryPool*.  Here: RunnerState* (server.cpp) = lgn_ctx handle + current mode,
    // iteration and pipe.  This is the one semantic change a ported runner has to make.
    void*        memorypool;
    // reference: GPUCache*, GPUGraphStorage*, GPUNodeStorage*.  Unused: the cache tiers, the CSR and the
    // seed sets are bound into the lgn_ctx (lgn_bind_*), not passed per call.
    void*        cache;
    void*        graph;
    void*        noder;
    // reference: IPCEnv*.  Same role here (compat IPCEnv wraps lgn_ipc_server_*).
    void*        env;
    // Fan-out of the hop a sampler operator draws (25 / 10 in the reference, Server.cu:68-69).
    int          neighbor_count;
    // true while presampling: samplers also count topology accesses, the planner counts feature accesses.
    bool         is_presc;
    // Host feature tier is mapped memory (always true on this side; the reference's SSD tier is out of scope).
    bool         in_memory;
};

class Operator {
public:
    virtual ~Operator() = default;
    // Enqueue this operator's work for the batch described by *params.  Asynchronous w.r.t. the host.
    virtual void run(OpParams* params) = 0;
};

// op_id follows the reference's numbering of the two-hop DAG (Server.cu:198-207):
//   0 batch generator | 1,3,5 feature extractor of segment (op_id-1)/2 | 2,4 sampler of hop op_id/2-1
//   6 cache planner (hotness accounting) | 7 cache updater (batch hand-off)
Operator* NewBatchGenerator(int op_id);
Operator* NewRandomSampler(int op_id);
Operator* NewFeatureExtractor(int op_id);
Operator* NewCachePlanner(int op_id);
Operator* NewCacheUpdater(int op_id);

#endif  // LEGION_B200_COMPAT_OPERATOR_H
