#!/bin/bash
# Accumulator ring depth (LRPX_TC_NBUF = 2 | 4) on the layers whose tile accumulators fit 128 TMEM columns.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -3
for nb in 2 4; do
  echo "nbuf=$nb $(LRPX_TC_NBUF=$nb LAYERS=0,1,2 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/nbuf_exp.log
for dbg in 1 4 5; do
  echo "nbuf=4 debug=$dbg $(LRPX_TC_NBUF=4 LRPX_TC_DEBUG=$dbg LAYERS=0 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror')"
done 2>&1 | tee -a gpurun_out/nbuf_exp.log
echo "nbuf=4 walk=1 $(LRPX_TC_NBUF=4 LRPX_TC_WALK=1 LAYERS=0 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror')" | tee -a gpurun_out/nbuf_exp.log
