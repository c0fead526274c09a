#!/bin/bash
# batches in flight vs hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS defaults to 8; a lane uses two streams)
cfg=${1:-C3}
for conn in 8 32; do for lanes in 4 6 8; do
  echo "== $cfg connections $conn lanes $lanes $V"
  env CUDA_DEVICE_MAX_CONNECTIONS=$conn $V python bench.py --config $cfg --probe --lanes $lanes 2>&1 | grep -E "only, $lanes|full"
done; done
