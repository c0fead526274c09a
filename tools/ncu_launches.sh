#!/bin/bash
# per-launch device times of a few steady-state steps (1 GPU).  usage: bash tools/ncu_launches.sh C2|C3 <tag>
CFG=${1:-C2}; TAG=${2:-r02}
OUT=gpurun_out
export LGN_BENCH_LONG_STEPS=0 LGN_BENCH_PRESAMPLE_STEPS=4
CMD="python bench.py --config $CFG --steps 2 --warmup 3 --no-cpu-baseline --no-train-epoch --no-parity"
$CMD > $OUT/ncu_plain_$TAG.log 2>&1 &&
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(sample|mark|assign|gather|batch)' -s 40 -c 40 --csv --log-file $OUT/${TAG}_launches_$CFG.csv $CMD > $OUT/ncu_$TAG.log 2>&1
echo "launch list $CFG rc=$?"
python - $OUT/${TAG}_launches_$CFG.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
seq = [(r[ki].split('(')[0][:30], float(r[vi].replace(',', '')) / 1000, r[gi]) for r in rows[hdr + 1:] if len(r) > vi and r[vi]]
idx = [i for i, s in enumerate(seq) if 'k_batch_begin' in s[0]]
for s in seq[idx[1]:idx[1] + 10]:
    print("%-32s %8.2f us grid %s" % s)
PY
