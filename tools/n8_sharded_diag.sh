#!/bin/bash
# Diagnostics for the open defect of DESIGN.md section 4 ("NVLink tier"): --placement sharded collapses at 8 GPUs
# once the shards are multi-GB.  One gpurun call:   gpurun --gpus 8 --timeout 600 -- 'bash tools/n8_sharded_diag.sh 8'
# Prints one line per experiment: tag, ms/step, alone gather launch (us), hit-mix fraction, payload GB/s per GPU.
N=${1:-8}
NODES=${2:-40000000}
OUT=gpurun_out
mkdir -p $OUT
RANKS=$(seq 0 $((N - 1)))

probe() {   # $1 = tag, rest = env assignments
    tag=$1; shift
    d=$(mktemp -d)
    for r in $RANKS; do env "$@" timeout 120 legion-1_b200/_build/peer_probe ipc $r $N $d > $OUT/probe_${tag}_$r.txt 2>&1 & done
    wait
    echo "== probe $tag (rank 0)"; cat $OUT/probe_${tag}_0.txt
}

bench() {   # $1 = tag, rest = env assignments
    tag=$1; shift
    env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 30 --warmup 5 --placement sharded --no-cpu-baseline --no-train-epoch --config C3 --nodes $NODES \
        2> $OUT/diag_$tag.err | tail -1 > $OUT/diag_$tag.json
    python - $OUT/diag_$tag.json $tag <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print("== bench %-14s %8.3f ms/step  alone %7.0f us  hit-mix %.3f  %6.0f GB/s/GPU" % (
        sys.argv[2], d["ms_per_step"], r["alone"]["avg_launch_us"], r["hit_mix"]["frac"], r["hit_mix"]["achieved_payload_GBps_per_gpu"]))
    if "peer_debug" in d.get("extra", {}):
        print("   in-process plain shard read, GB/s per GPU:", d["extra"]["peer_debug"])
except Exception as e:
    print("== bench %s failed: %r" % (sys.argv[2], e))
PY
}

bench control LGN_BENCH_PEER_DEBUG=1     # reproduces the defect; extra.peer_debug = the probe's plain loop over THIS process's mappings
bench gloo LGN_BENCH_BACKEND=gloo LGN_BENCH_PEER_DEBUG=1   # no NCCL communicator in the process at all
bench vmm LGN_BENCH_SHARD_ALLOC=vmm LGN_BENCH_PEER_DEBUG=1  # shards through cuMemCreate + fd + cuMemSetAccess instead of cudaIpc*
bench nonvls NCCL_NVLS_ENABLE=0          # NCCL without NVLink SHARP multicast
bench nohot LGN_BENCH_HOTNESS=none       # NCCL communicator, but no large all-reduce
bench hosthot LGN_BENCH_HOTNESS=host     # hotness reduced through host memory
probe alloc8g A=1                        # 8 GiB tables + product-like variants (hints / skew / lookup)
probe alloc2g5 PROBE_ALLOC_MB=2441       # the product's 2.56 GB shard size
