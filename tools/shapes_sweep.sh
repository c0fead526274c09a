#!/bin/bash
# One bench line per BASELINE.json shape that fits a single B200 at reduced node count (parity for these shapes is
# covered by tests/; this is the measurement side).   gpurun --timeout 900 -- 'bash tools/shapes_sweep.sh'
OUT=gpurun_out
mkdir -p $OUT
run() {   # $1 = tag, rest = bench arguments
    tag=$1; shift
    timeout 400 python bench.py --steps 50 --warmup 10 --no-cpu-baseline "$@" 2> $OUT/shape_$tag.err | tail -1 > $OUT/shape_$tag.json
    python - $OUT/shape_$tag.json $tag <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]; x = d["extra"]
    print("== %-22s %7.3f ms/step %6.2f G edges/s  %6.0f GB/s features  frac %.2f hit-mix %.2f tiers %s  epoch %s s" % (
        sys.argv[2], d["ms_per_step"], d["value"] / 1e9, x["feature_extract_GBps"], r["frac"], r["hit_mix"]["frac"], x["tier_rows"],
        x.get("graphsage_epoch_s")))
except Exception as e:
    print("== %s failed: %r" % (sys.argv[2], e))
PY
}
run C2_products                    --config C2
run C3_papers100M_full             --config C3 --no-train-epoch
run C4_uk_host_tier_20M_nodes      --config C4 --nodes 20000000 --cache-frac 0.3 --no-train-epoch      # 70 % of the rows served from pinned host memory over UVA
run C5_friendster_3hop_16M_nodes   --config C5 --nodes 16000000                                        # fanout [15,10,5]
