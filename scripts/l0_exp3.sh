#!/bin/bash
mkdir -p gpurun_out
for cfg in "X=1" "LRPX_TC_MH=1" "LRPX_TC_ISSUERS=1" "LRPX_TC_DEBUG=16" "LRPX_TC_DEBUG=18" "LRPX_TC_DEBUG=20" "LRPX_TC_DEBUG=22" "LRPX_TC_DEBUG=7"; do
  echo "$cfg $(env $cfg LAYERS=0 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | sed 's/ max [0-9.]* ms//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/l0_exp3.log
