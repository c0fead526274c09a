#!/bin/bash
# ncu evidence of the current kernels (1 GPU).  gpurun --timeout 900 -- 'bash tools/ncu_r02.sh'
OUT=gpurun_out
mkdir -p $OUT
export LGN_BENCH_LONG_STEPS=0
C2="python bench.py --config C2 --steps 4 --warmup 3 --no-cpu-baseline --no-train-epoch --no-parity"
C3="python bench.py --config C3 --steps 4 --warmup 3 --no-cpu-baseline --no-train-epoch --no-parity"
$C2 > $OUT/ncu_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(sample|mark|assign|gather|batch)' -s 245 -c 120 --csv --log-file $OUT/r02_launches_c2.csv $C2 > $OUT/ncu_c2.log 2>&1
echo "launch list c2 rc=$?"
$C3 > $OUT/ncu_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_(sample|mark|assign|gather|batch)' -s 11130 -c 22 -o $OUT/r02_prof_c3 $C3 > $OUT/ncu_c3.log 2>&1
echo "set full c3 rc=$?"
ls -la $OUT/*.ncu-rep
