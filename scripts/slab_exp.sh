#!/bin/bash
for dbg in 0 1 3 5 7 17 19 23; do
  echo "== LRPX_TC_DEBUG=$dbg (1 skip epi io, 2 skip mma, 4 skip A loads, 16 skip epilogue tmem ld)"
  LRPX_TC_SLAB=1 LRPX_TC_DEBUG=$dbg LAYERS="${LAYERS:-0,1,3}" REPS=15 python scripts/one_layer.py 2>&1 | grep layer
done
