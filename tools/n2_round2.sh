#!/bin/bash
# 2-GPU round-2 call: gather variant per placement, ncu link counters of the gather (single process, two GPUs), the
# server binary with the unchanged trainer.   gpurun --gpus 2 --timeout 900 -- 'bash tools/n2_round2.sh'
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for g in bulk ldg; do
  LGN_GATHER=$g LGN_BENCH_LONG_STEPS=0 timeout 300 $TR bench.py --gpus 2 --no-train-epoch > $OUT/r2k_n2_$g.json 2> $OUT/r2k_n2_$g.err
  python - $OUT/r2k_n2_$g.json $g <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); s = d["extra"]["sharded"]
print("== gather %-4s hybrid %.4f ms/step (peer %.3f, hit-mix %.3f)   sharded %.4f ms/step (hit-mix %.3f)  clocks %s" % (
    sys.argv[2], d["ms_per_step"], d["roofline"]["hit_mix"]["peer"], d["roofline"]["hit_mix"]["frac"], s["ms_per_step"], s["hit_mix"]["frac"], d["clocks"]))
PY
done
M=nvlrx__bytes.sum,nvltx__bytes.sum,pcie__read_bytes.sum,pcie__write_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors.sum,gpu__time_duration.sum
python tools/nvlink_ncu.py > $OUT/nvl_plain.log 2>&1 && cat $OUT/nvl_plain.log | tail -1 &&
timeout 300 ncu --metrics $M -k regex:k_gather --clock-control none --csv --log-file $OUT/r02k_nvlink_gather.csv python tools/nvlink_ncu.py > $OUT/nvl_ncu.log 2>&1
echo "ncu nvlink rc=$?"
NVL_HOST_FRAC=0.5 python tools/nvlink_ncu.py > $OUT/nvl_host_plain.log 2>&1 && cat $OUT/nvl_host_plain.log | tail -1 &&
NVL_HOST_FRAC=0.5 timeout 300 ncu --metrics $M -k regex:k_gather --clock-control none --csv --log-file $OUT/r02k_nvlink_host_gather.csv python tools/nvlink_ncu.py > $OUT/nvl_host_ncu.log 2>&1
echo "ncu nvlink+host rc=$?"
timeout 300 python tools/server_e2e.py --gpus 2 --config C2 --agg-mode 1 --trainer legion_graphsage --epochs 3 > $OUT/r2k_e2e_sage_n2.json 2> $OUT/r2k_e2e_sage_n2.err
echo "e2e sage n2 rc=$?"; cut -c1-600 $OUT/r2k_e2e_sage_n2.json
