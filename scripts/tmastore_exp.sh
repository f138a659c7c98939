#!/bin/bash
for v in 1 0; do
  echo "== LRPX_TC_TMASTORE=$v"
  LRPX_TC_TMASTORE=$v LAYERS="${LAYERS:-1,3,5,6,8,9,11,12}" REPS=9 python scripts/one_layer.py 2>&1 | grep layer
done
