mkdir -p gpurun_out
PCMD="python bench.py --profile-step --no-graph"
timeout 900 $PCMD > gpurun_out/plain2.log 2>&1 &&
timeout 1800 ncu --set full --clock-control none -k regex:tc_conv_slab -s 82 -c 13 -o gpurun_out/prof_chain $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log
