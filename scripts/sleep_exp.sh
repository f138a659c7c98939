#!/bin/bash
# Back-off of the relaxed barrier waits (LRPX_TC_SLEEP, ns) per high-resolution chain layer, row walk on/off for layer 0.
mkdir -p gpurun_out
for sl in 200 100 50 20 0; do
  echo "sleep=$sl $(LRPX_TC_SLEEP=$sl LAYERS=0,1,2,3,4,6,9 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | tr '\n' '|')"
done 2>&1 | tee gpurun_out/sleep_exp.log
for sl in 200 20 0; do
  echo "walk=0 sleep=$sl $(LRPX_TC_WALK=0 LRPX_TC_SLEEP=$sl LAYERS=0 REPS=9 timeout 100 python scripts/one_layer.py 2>&1 | grep 'layer\|rror')"
done 2>&1 | tee -a gpurun_out/sleep_exp.log
