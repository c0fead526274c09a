#!/bin/bash
# whole-range counters under real concurrency (4 lanes, graph replay off): ncu --replay-mode app-range around 40 steps.
# usage: bash tools/ncu_range.sh C2|C3 samp|both <tag>
CFG=${1:-C2}; WHAT=${2:-both}; TAG=${3:-r02}
OUT=gpurun_out
export LGN_BENCH_PRESAMPLE_STEPS=8 LGN_GRAPH=0 LGN_NCU_RANGE=$WHAT
CMD="python bench.py --config $CFG --probe --no-cpu-baseline"
$CMD > $OUT/range_plain_${TAG}_${CFG}_$WHAT.log 2>&1 &&
timeout 400 ncu --replay-mode app-range --clock-control none --cache-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    --csv --log-file $OUT/${TAG}_range_${CFG}_$WHAT.csv $CMD > $OUT/range_ncu_${TAG}_${CFG}_$WHAT.log 2>&1
echo "range $CFG $WHAT rc=$?"
grep -E "dram__|lts__|gpu__time" $OUT/${TAG}_range_${CFG}_$WHAT.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
