#!/bin/bash
# grid caps of the sampling kernels and the gather variant, timed over 400 pipelined steps (probe)
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" python bench.py --config $cfg --probe 2>&1 | grep -E "only, 4|full"; }
run C3 LGN_X=0
run C3 LGN_GATHER=bulk LGN_CARVEOUT=50
run C3 LGN_GATHER=bulk LGN_CARVEOUT=50 LGN_GATHER_THREADS=128
run C3 LGN_RESOLVE_CTAS=6
run C3 LGN_RESOLVE_CTAS=16
run C3 LGN_SAMPLE_CTAS=8
run C3 LGN_END_CTAS=2
run C3 LGN_GATHER_LDG_CTAS=3
run C3 LGN_GATHER_LDG_CTAS=1 LGN_GATHER_UNROLL=8
run C3 LGN_GATHER_LDG_CTAS=2 LGN_GATHER_UNROLL=8
run C2 LGN_X=0
run C2 LGN_GATHER=bulk
run C2 LGN_GATHER_LDG_CTAS=3
