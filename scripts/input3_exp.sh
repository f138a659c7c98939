#!/bin/bash
# folded first-layer kernel (LRPX_TC_INPUT3) vs the N=16 form: parity tests, then the chain's layer-0 time via bench
for v in 1 0; do
  echo "== LRPX_TC_INPUT3=$v"
  LRPX_TC_INPUT3=$v timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "engine or relevance_groups" 2>&1 | grep -v "mbarrier wait" | tail -3
done
