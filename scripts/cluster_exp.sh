#!/bin/bash
# CTA-pair multicast of the streamed B tiles (LRPX_TC_CLUSTER=1) vs off: parity tests, then per-layer times
LRPX_TC_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x 2>&1 | grep -v "mbarrier wait" | tail -4
for v in 1 0; do
  echo "== LRPX_TC_CLUSTER=$v"
  LRPX_TC_CLUSTER=$v LAYERS="${LAYERS:-3,4,5,7,9,10,12}" REPS=7 timeout 120 python scripts/one_layer.py 2>&1 | grep "layer\|Error\|error" | head -12
done
