// A GPURunner in the reference's style (Server.cu:169-335, Operator.cu:10-123), written against the reference's entry
// points only: storages from the factories, a GPUMemoryPool that owns its own d_alloc_space buffers, the five kernel
// entry points in the reference's op order, GPUCache::CandidateSelection / CostModel / FillUp after a presampling pass.
// Reads a dataset in the reference's on-disk format, writes every batch's buffers to <out>; tests/test_reference_abi.py
// compares them with the CPU oracle.   usage: ref_style_runner <dataset dir> <N> <E> <D> <batch> <n_batches> <cache bytes> <out>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "reference_abi.h"

static void slurp(const std::string& path, void* dst, size_t bytes)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f || fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
    fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 9) return 2;
    const std::string dir = argv[1];
    const int32_t N = atoi(argv[2]), D = atoi(argv[4]), B = atoi(argv[5]), n_batches = atoi(argv[6]);
    const int64_t E = atoll(argv[3]), cache_bytes = atoll(argv[7]);
    const int f1 = 25, f2 = 10, dev = 0, depth = 2;
    SetGPUDevice(dev);
    // GPUGraphStore::Load_Graph / Load_Feature: pinned + mapped host arrays (GPUGraphStore.cu:254-427)
    std::vector<int64_t> indptr(N + 1);
    std::vector<int32_t> indices(E), labels(N);
    std::vector<float> feats((size_t)N * D);
    slurp(dir + "/edge_src", indptr.data(), (size_t)(N + 1) * 8);
    slurp(dir + "/edge_dst", indices.data(), (size_t)E * 4);
    slurp(dir + "/features", feats.data(), (size_t)N * D * 4);
    slurp(dir + "/labels", labels.data(), (size_t)N * 4);
    BuildInfo info;
    info.partition_count = 1; info.shard_to_partition = {0}; info.shard_to_device = {dev};
    info.total_num_nodes = N; info.float_attr_len = D; info.total_edge_num = E; info.raw_batch_size = B; info.epoch = 1;
    info.csr_node_index = (int64_t*)host_alloc_space((unsigned)((N + 1) * 8));
    info.csr_dst_node_ids = (int32_t*)host_alloc_space((unsigned)(E * 4));
    info.host_float_attrs = (float*)host_alloc_space((unsigned)((size_t)N * D * 4));
    cudaMemcpy(info.csr_node_index, indptr.data(), (size_t)(N + 1) * 8, cudaMemcpyDefault);
    cudaMemcpy(info.csr_dst_node_ids, indices.data(), (size_t)E * 4, cudaMemcpyDefault);
    cudaMemcpy(info.host_float_attrs, feats.data(), (size_t)N * D * 4, cudaMemcpyDefault);
    std::vector<int32_t> train, train_lab;
    for (int32_t i = 0; i < N; i += 3) { train.push_back(i); train_lab.push_back(labels[i]); }     // every third node is a seed
    info.training_set_num = {(int32_t)train.size()}; info.training_set_ids = {train}; info.training_labels = {train_lab};
    info.validation_set_num = {0}; info.validation_set_ids = {{}}; info.validation_labels = {{}};
    info.testing_set_num = {0}; info.testing_set_ids = {{}}; info.testing_labels = {{}};
    GPUGraphStorage* graph = NewGPUMemoryGraphStorage();
    GPUNodeStorage* noder = NewGPUMemoryNodeStorage();
    graph->Build(&info);
    noder->Build(&info);
    const int32_t train_step = ((int32_t)train.size() - 1) / B;
    GPUCache cache;
    cache.Initialize(cache_bytes, 0, D, train_step, 1);
    cache.InitializeCacheController(dev, N);
    // GPURunner::Initialize (Server.cu:217-283): the runner allocates the wire buffers and registers them in its pool
    const int64_t num_ids = (int64_t)B * (1 + f1 + f1 * f2);
    GPUMemoryPool pool(depth);
    for (int p = 0; p < depth; p++) {
        pool.SetSampledIds((int32_t*)d_alloc_space(num_ids * 4), p);
        pool.SetLabels((int32_t*)d_alloc_space((int64_t)B * 4), p);
        pool.SetAggSrcOf((int32_t*)d_alloc_space(num_ids * 4), p);
        pool.SetAggDstOf((int32_t*)d_alloc_space(num_ids * 4), p);
        pool.SetNodeCounter((int32_t*)d_alloc_space(64), p);
        pool.SetEdgeCounter((int32_t*)d_alloc_space(64), p);
        pool.SetFloatFeatures((float*)d_alloc_space(num_ids * D * 4), p);
    }
    pool.SetBufferSizes(num_ids, num_ids);
    cudaStream_t s0, s1;
    cudaStreamCreate(&s0);
    cudaStreamCreate(&s1);
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    auto run = [&](int iter, bool presc) {      // the operator DAG of RunPreSc / RunOnce (Server.cu:284-328)
        pool.SetCurrentPipe(iter % depth);
        pool.SetCurrentMode(0);
        pool.SetIter(iter);
        batch_generator_kernel(s0, noder, &cache, &pool, B, iter, dev, dev, 0);
        if (!presc) { cudaEventRecord(ev, s0); cudaStreamWaitEvent(s1, ev, 0); get_feature_kernel(s1, &cache, noder, &pool, dev, 1, true); }
        GPU_Random_Sampling(s0, graph, &cache, &pool, f1, 2, presc);
        if (!presc) { cudaEventRecord(ev, s0); cudaStreamWaitEvent(s1, ev, 0); get_feature_kernel(s1, &cache, noder, &pool, dev, 3, true); }
        GPU_Random_Sampling(s0, graph, &cache, &pool, f2, 4, presc);
        if (!presc) { cudaEventRecord(ev, s0); cudaStreamWaitEvent(s1, ev, 0); get_feature_kernel(s1, &cache, noder, &pool, dev, 5, true); }
        make_update_plan(s0, graph, &cache, &pool, dev, 0);
        update_cache(s0, &cache, noder, &pool, dev, 0);
        cudaStreamSynchronize(s0);
        cudaStreamSynchronize(s1);
    };
    for (int it = 0; it < train_step; it++) run(it, true);                       // GPUServer::PreSc (Server.cu:83-114)
    std::vector<uint64_t> counters;                                              // no PCM here: the layer counts transactions itself
    cache.CandidateSelection(0, noder, graph);
    cache.CostModel(0, noder, graph, counters, train_step);
    cache.FillUp(0, noder, graph);
    FILE* out = fopen(argv[8], "wb");
    int32_t hdr[4] = {n_batches, cache.NodeCapacity(dev), cache.MaxIdNum(dev), train_step};
    fwrite(hdr, 4, 4, out);
    for (int it = 0; it < n_batches; it++) {
        run(it, false);
        int32_t nc[16], ec[16];
        d_copy_2_h(nc, pool.GetNodeCounter(), 64);
        d_copy_2_h(ec, pool.GetEdgeCounter(), 64);
        const int32_t total = nc[9], n_e = ec[4];
        std::vector<int32_t> ids(total), lab(nc[4]), so(n_e), dof(n_e);
        std::vector<float> ft((size_t)total * D);
        d_copy_2_h(ids.data(), pool.GetSampledIds(), total * 4);
        d_copy_2_h(lab.data(), pool.GetLabels(), nc[4] * 4);
        d_copy_2_h(so.data(), pool.GetAggSrcOf(), n_e * 4);
        d_copy_2_h(dof.data(), pool.GetAggDstOf(), n_e * 4);
        d_copy_2_h(ft.data(), pool.GetFloatFeatures(), (unsigned)((size_t)total * D * 4));
        fwrite(nc, 4, 16, out); fwrite(ec, 4, 16, out);
        fwrite(ids.data(), 4, total, out); fwrite(lab.data(), 4, nc[4], out);
        fwrite(so.data(), 4, n_e, out); fwrite(dof.data(), 4, n_e, out);
        fwrite(ft.data(), 4, (size_t)total * D, out);
    }
    fclose(out);
    printf("ok %d batches, node capacity %d\n", n_batches, hdr[1]);
    return 0;
}
