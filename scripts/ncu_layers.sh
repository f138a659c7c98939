#!/bin/bash
# ncu --set full on single chain layers in both A-fetch modes -> gpurun_out/prof_L_<mode>.ncu-rep
mkdir -p gpurun_out
for mode in 0 1; do
  export LRPX_TC_SLAB=$mode
  LAYERS="${LAYERS:-1,3,9}" REPS=2 python scripts/one_layer.py > gpurun_out/one_layer_$mode.log 2>&1 &&
  LAYERS="${LAYERS:-1,3,9}" REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 12 -c 6 \
     -o gpurun_out/prof_layers_slab$mode python scripts/one_layer.py > gpurun_out/ncu_layers_$mode.log 2>&1
  echo "mode $mode ncu exit $?"; cat gpurun_out/one_layer_$mode.log
done
