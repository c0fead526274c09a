for cfg in C2 C3; do
for v in "LGN_SAMPLE_CTAS=8" "LGN_SAMPLE_CTAS=16" "LGN_RESOLVE_CTAS=16" "LGN_RESOLVE_CTAS=6" "LGN_GATHER_CTAS=2" "LGN_END_CTAS=8"; do
  echo "== $cfg $v"; env $v python bench.py --config $cfg --probe 2>&1 | grep -E "only, 4|full"
done
echo "== $cfg lanes 6"; python bench.py --config $cfg --probe --lanes 6 2>&1 | grep -E "full"
echo "== $cfg lanes 8"; python bench.py --config $cfg --probe --lanes 8 2>&1 | grep -E "only, 8|full"
done
