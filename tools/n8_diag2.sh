#!/bin/bash
# Round-2 diagnostic of the 8-GPU partitioned-cache collapse: one process per GPU, bench.py --probe
# (in-process plain shard read through the real peer mappings, sampling-only / gather-only / full loops), with NVLink
# byte counters around the first run.   gpurun --gpus 8 --timeout 420 -- 'bash tools/n8_diag2.sh 8 40000000 control gloo'
N=${1:-8}
NODES=${2:-40000000}
shift 2
OUT=gpurun_out
mkdir -p $OUT
run() {
    tag=$1; shift
    env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --probe --placement sharded --no-cpu-baseline --no-train-epoch --config C3 --nodes $NODES \
        > $OUT/d2_$tag.out 2> $OUT/d2_$tag.err
    echo "== $tag rc=$?"
    grep -E "peer_debug|sampling only|gather only|full pipeline" $OUT/d2_$tag.err | sort | uniq -c | sort -k2 | head -60
}
nvidia-smi nvlink -gt d -i 0 > $OUT/d2_nvlink_before.txt 2>&1
nvidia-smi topo -p2p r > $OUT/d2_p2p_r.txt 2>&1
for tag in "$@"; do
    case $tag in
        control) run control LGN_BENCH_PEER_DEBUG=1 ;;
        gloo) run gloo LGN_BENCH_BACKEND=gloo LGN_BENCH_PEER_DEBUG=1 ;;
        nohot) run nohot LGN_BENCH_HOTNESS=none LGN_BENCH_PEER_DEBUG=1 ;;
        nonvls) run nonvls NCCL_NVLS_ENABLE=0 LGN_BENCH_PEER_DEBUG=1 ;;
        vmm) run vmm LGN_BENCH_SHARD_ALLOC=vmm LGN_BENCH_PEER_DEBUG=1 ;;
    esac
    nvidia-smi nvlink -gt d -i 0 > $OUT/d2_nvlink_after_$tag.txt 2>&1
done
