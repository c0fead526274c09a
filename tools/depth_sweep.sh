#!/bin/bash
# bytes the gather keeps in flight per SM vs the step time with the sampling chains beside it
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" python bench.py --config $cfg --probe 2>&1 | grep -E "gather only|full"; }
cfg=${1:-C3}
run $cfg LGN_GATHER_THREADS=64
run $cfg LGN_GATHER_THREADS=96
run $cfg LGN_GATHER_THREADS=128
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=1
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=1 LGN_GATHER_UNROLL=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=2 LGN_GATHER_UNROLL=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=2 LGN_CARVEOUT=30
