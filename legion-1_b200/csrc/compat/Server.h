// Server.h -- source-compatible with the reference server API (reference Server.h:137-165):
// class Server {Initialize, PreSc, Run, Finalize} + NewGPUServer(), class Runner + NewGPURunner(),
// struct RunnerParams.  main.cpp is the reference's main.cpp:4-10 call sequence.
#ifndef LEGION_B200_SERVER_H
#define LEGION_B200_SERVER_H
#include <vector>

struct RunnerParams {
    int device_id;
    std::vector<int> fanout;
    void* cache;
    void* graph;
    void* noder;
    void* env;
    int global_batch_id;
    bool in_memory;
};

class Server {
public:
    virtual ~Server() = default;
    virtual void Initialize(int global_shard_count) = 0;
    virtual void PreSc(int cache_agg_mode) = 0;
    virtual void Run() = 0;
    virtual void Finalize() = 0;
};
Server* NewGPUServer();

class Runner {
public:
    virtual ~Runner() = default;
    virtual void Initialize(RunnerParams* params) = 0;
    virtual void InitializeFeaturesBuffer(RunnerParams* params) = 0;
    virtual void RunPreSc(RunnerParams* params) = 0;
    virtual void RunOnce(RunnerParams* params) = 0;
    virtual void Finalize(RunnerParams* params) = 0;
};
Runner* NewGPURunner();

#endif
