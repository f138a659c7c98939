#!/bin/bash
for dbg in 0 32 64 96 1; do
  echo "== LRPX_TC_DEBUG=$dbg (32 no gain loads, 64 no stores, 1 no epilogue global traffic)"
  LRPX_TC_DEBUG=$dbg LAYERS="${LAYERS:-1,3,5}" REPS=9 python scripts/one_layer.py 2>&1 | grep layer
done
