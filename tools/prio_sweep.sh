#!/bin/bash
# stream-priority / gather-residency sweep of the probe: which kernel gets SM slots first when the
# sampling chains of the other lanes and a feature gather are both pending
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" python bench.py --config $cfg --probe 2>&1 | grep -E "only, 4|only, 1 streams|full"; }
for cfg in C3 C2; do
  run $cfg LGN_X=0
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=lo
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=hi
  run $cfg LGN_GATHER_PRIO=lo LGN_LANE_PRIO=lo
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=lo LGN_GATHER_CTAS=2
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=hi LGN_GATHER_CTAS=2
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=lo LGN_GRAPH=0
  run $cfg LGN_GATHER_PRIO=hi LGN_LANE_PRIO=lo LGN_SAMPLE_CTAS=4 LGN_RESOLVE_CTAS=4
done
