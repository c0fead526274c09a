#!/bin/bash
# round-2 8-GPU measurement: the default bench line (hybrid + extra.sharded, papers100M shape) and the server binary with
# the reference's unchanged GraphSAGE trainer.   gpurun --gpus 8 --timeout 600 -- 'bash tools/n8_final.sh 8'
N=${1:-8}
OUT=gpurun_out
TAG=${2:-r2j}
mkdir -p $OUT
date +%s
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    > $OUT/${TAG}_n$N.json 2> $OUT/${TAG}_n$N.err
echo "bench n$N rc=$?"; date +%s
tail -c 600 $OUT/${TAG}_n$N.err
MODE=$(python -c "print({1:0,2:1,4:2,8:3}[$N])")
timeout 300 python tools/server_e2e.py --gpus $N --config C2 --agg-mode $MODE --trainer legion_graphsage --epochs 3 \
    > $OUT/${TAG}_e2e_sage_n$N.json 2> $OUT/${TAG}_e2e_sage_n$N.err
echo "e2e sage n$N rc=$?"; date +%s
cut -c1-700 $OUT/${TAG}_e2e_sage_n$N.json; tail -c 400 $OUT/${TAG}_e2e_sage_n$N.err
free -g | head -2
