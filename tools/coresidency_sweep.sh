#!/bin/bash
# how much of each SM the gather may hold while the sampling chains of the other lanes run beside it
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" python bench.py --config $cfg --probe $EXTRA 2>&1 | grep -E "only, 4|full"; }
cfg=${1:-C3}
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=4
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3 LGN_GATHER_UNROLL=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=4 LGN_GATHER_UNROLL=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=6 LGN_GATHER_UNROLL=2
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3 LGN_GATHER_PRIO=hi LGN_LANE_PRIO=lo
EXTRA="--lanes 6" run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3
run $cfg LGN_CARVEOUT=60 LGN_GATHER_THREADS=128
run $cfg LGN_CARVEOUT=50
run $cfg LGN_CARVEOUT=70
run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3 LGN_CARVEOUT=30
