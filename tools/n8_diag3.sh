#!/bin/bash
# Round-2 diagnostic, part 2: is the reach of the address translation (page size of the shard mappings) what bounds random
# reads out of multi-GB peer shards?   gpurun --gpus 8 --timeout 420 -- 'bash tools/n8_diag3.sh 8 40000000'
N=${1:-8}
NODES=${2:-40000000}
OUT=gpurun_out
mkdir -p $OUT
RANKS=$(seq 0 $((N - 1)))
probe() {   # $1 = tag, rest = env assignments
    tag=$1; shift
    d=$(mktemp -d)
    for r in $RANKS; do env PROBE_KINDS=1 "$@" timeout 100 legion-1_b200/_build/peer_probe ipc $r $N $d > $OUT/p3_${tag}_$r.txt 2>&1 & done
    wait
    echo "== probe $tag (rank 0)"; cat $OUT/p3_${tag}_0.txt
}
bench() {   # $1 = tag, rest = env assignments
    tag=$1; shift
    env LGN_BENCH_PEER_DEBUG=exit "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --placement sharded --no-cpu-baseline --no-train-epoch --config C3 --nodes $NODES \
        > $OUT/d3_$tag.out 2> $OUT/d3_$tag.err
    echo "== bench $tag rc=$?"
    grep -E "peer_debug" $OUT/d3_$tag.err | sort | head -3
}
probe a2441 PROBE_ALLOC_MB=2441
probe a2560 PROBE_ALLOC_MB=2560
probe a2441_pre25g PROBE_ALLOC_MB=2441 PROBE_PREALLOC_MB=25000
probe a7168 PROBE_ALLOC_MB=7168
bench vmm512 LGN_BENCH_SHARD_ALLOC=vmm
bench vmm2 LGN_BENCH_SHARD_ALLOC=vmm LGN_VMM_ALIGN_MB=2
bench ipcpad512 LGN_BENCH_SHARD_PAD_MB=512
bench ipcearly LGN_BENCH_EARLY_SHARD=1
