#!/bin/bash
# first-layer relevance kernel: parity tests, then its time (folded columns vs the N=16 form)
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "folded or engine" 2>&1 | grep -v "mbarrier wait" | tail -3
for v in 1 0; do
  echo "== LRPX_TC_INPUT3=$v"
  LRPX_TC_INPUT3=$v LAYERS=0 REPS=9 timeout 100 python scripts/one_layer.py 2>&1 | grep layer
done
