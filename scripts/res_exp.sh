#!/bin/bash
for cfg in "LRPX_TC_RES3=1" "LRPX_TC_RES3=0"; do
  echo "== $cfg"
  env $cfg LAYERS="${LAYERS:-0,1,2,3,4}" REPS=9 python scripts/one_layer.py 2>&1 | grep layer
done
