#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -3
for cfg in "X=1" "LRPX_TC_AGROUP=0" "LRPX_TC_NBUF=2" "LRPX_TC_DEBUG=16" "LRPX_TC_DEBUG=22" "LRPX_TC_DEBUG=1" "LRPX_TC_DEBUG=4"; do
  echo "$cfg $(env $cfg LAYERS=0 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | sed 's/ max [0-9.]* ms//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/l0_exp4.log
