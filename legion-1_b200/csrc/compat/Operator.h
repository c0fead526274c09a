// Operator.h -- source-compatible with the reference plugin boundary (reference Operator.h:4-27):
// same struct, same abstract class, same five factories.  The implementations (operator.cpp) are
// thin calls into the C-ABI (include/legion_b200.h); `memorypool` carries the per-GPU RunnerState.
#ifndef LEGION_B200_OPERATOR_H
#define LEGION_B200_OPERATOR_H
#include <cuda_runtime.h>

struct OpParams {
    int device_id;
    cudaStream_t stream;
    cudaEvent_t event;
    void* memorypool;   // RunnerState* (reference: GPUMemoryPool*)
    void* cache;        // unused by the B200 operators: the caches live inside the lgn_ctx
    void* graph;
    void* noder;
    void* env;          // IPCEnv*
    int neighbor_count;
    bool is_presc;
    bool in_memory;
};

class Operator {
public:
    virtual ~Operator() = default;
    virtual void run(OpParams* params) = 0;
};

Operator* NewBatchGenerator(int op_id);
Operator* NewRandomSampler(int op_id);
Operator* NewFeatureExtractor(int op_id);
Operator* NewCachePlanner(int op_id);
Operator* NewCacheUpdater(int op_id);

#endif
