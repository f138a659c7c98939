#!/bin/bash
for dbg in 0 1 2 3 4 5 8 9 12 13; do
  echo "== LRPX_TC_DEBUG=$dbg (1 skip epi io, 2 skip mma, 4 skip A loads, 8 aligned A views)"
  LRPX_TC_SLAB=1 LRPX_TC_DEBUG=$dbg LAYERS="0,1,3,9" REPS=3 python scripts/one_layer.py 2>&1 | grep layer | tr '\n' ' '; echo
done
