#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcx.py -m gpu -x -q 2>&1 | tail -3
for cfg in "X=1" "LRPX_TC_NBUF=2" "LRPX_TC_NBUF=4" "LRPX_TC_WALK=1" "LRPX_TC_DEBUG=1" "LRPX_TC_DEBUG=4" "LRPX_TC_DEBUG=5"; do
  echo "$cfg $(env $cfg LAYERS=${LAYERS:-0,1,2,3,4} REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | sed 's/ max [0-9.]* ms//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/l0_exp2.log
