#!/bin/bash
mkdir -p gpurun_out
for cfg in "X=1" "LRPX_TC_PAIR=2" "LRPX_TC_PAIR=2 LRPX_TC_NBUF=4" "LRPX_TC_NBUF=4" "X=1"; do
  echo "$cfg $(env $cfg LAYERS=1,2 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | sed 's/ max [0-9.]* ms//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/pair64_exp.log
