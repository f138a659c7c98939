#!/bin/bash
# MMA issue-rate probe: relevance-chain layers with the epilogue's global traffic (1) and the A loads (4) switched off
for cfg in "LRPX_TC_MH=2 LRPX_TC_ISSUERS=2" "LRPX_TC_MH=2 LRPX_TC_ISSUERS=1" "LRPX_TC_MH=1 LRPX_TC_ISSUERS=1"; do
  for dbg in ${DBGS:-0 5 21}; do
    echo "== $cfg LRPX_TC_DEBUG=$dbg"
    env $cfg LRPX_TC_SLAB=1 LRPX_TC_DEBUG=$dbg LAYERS="${LAYERS:-0,1,2,4}" REPS=9 python scripts/one_layer.py 2>&1 | grep layer
  done
done
