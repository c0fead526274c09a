#!/bin/bash
# shared-memory carveout / gather-variant sweep of the probe: do the sampling kernels and the gather really co-reside on an SM?
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" python bench.py --config $cfg --probe 2>&1 | grep -E "only, 4|full"; }
for cfg in C3; do
  run $cfg LGN_X=0
  run $cfg LGN_CARVEOUT=100
  run $cfg LGN_CARVEOUT=60
  run $cfg LGN_GATHER=ldg
  run $cfg LGN_GATHER=ldg LGN_CARVEOUT=100
  run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3
  run $cfg LGN_GATHER=ldg LGN_GATHER_LDG_CTAS=3 LGN_CARVEOUT=0
done
