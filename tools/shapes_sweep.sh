#!/bin/bash
# One bench line per BASELINE.json shape that fits a single B200 (C4 / C5 at reduced node count); every line carries
# parity_checked = true (one batch against the CPU oracle before timing).   gpurun --timeout 900 -- 'bash tools/shapes_sweep.sh'
OUT=gpurun_out
mkdir -p $OUT
export LGN_BENCH_LONG_STEPS=0
run() {   # $1 = tag, rest = bench arguments
    tag=$1; shift
    timeout 400 python bench.py --steps 50 --warmup 10 --no-cpu-baseline "$@" 2> $OUT/shape_$tag.err | tail -1 > $OUT/shape_$tag.json
    python - $OUT/shape_$tag.json $tag <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]; x = d["extra"]; h = r["hit_mix"]
    print("== %-28s %7.3f ms/step %6.2f G edges/s %6.0f GB/s features  step frac %.2f  hit-mix %.2f (local %.2f peer %.2f host %.2f)  parity %s  epoch %s s" % (
        sys.argv[2], d["ms_per_step"], d["value"] / 1e9, x["feature_extract_GBps"], r["step"]["frac"], h["frac"], h["local"], h["peer"], h["host"],
        d["parity_checked"], x.get("graphsage_epoch_s")))
except Exception as e:
    print("== %s failed: %r" % (sys.argv[2], e))
PY
}
run C2_products                    --config C2
run C4_uk_20M_nodes_all_cached     --config C4 --nodes 20000000 --no-train-epoch
run C4_uk_20M_nodes_host_tier      --config C4 --nodes 20000000 --cache-frac 0.3 --no-train-epoch      # only the 30 % hottest rows may be cached: the rest over UVA
run C5_friendster_3hop_16M_nodes   --config C5 --nodes 16000000 --no-train-epoch                       # fanout [15,10,5]
LGN_GATHER=ldg run C4_uk_20M_host_tier_ldg_gather --config C4 --nodes 20000000 --cache-frac 0.3 --no-train-epoch
