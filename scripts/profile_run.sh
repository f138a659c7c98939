#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the default bench workload (one warm-up + one step), (2) --set full on
# the 13 relevance-chain layers of the measured step of that same command: per step the slab kernel is launched 12
# times by the forward (FWD_GAIN) and 58 times by the chain (8 low-resolution layers over all 1216 requests, then 10
# chunks x 5 high-resolution layers), so launches 82 .. 94 of tc_conv_slab_kernel are the 8 stage-1 launches and the
# 5 launches of the first chunk of the second step.
mkdir -p gpurun_out
PCMD="python bench.py --profile-step --no-graph"
timeout 900 $PCMD > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
[ -n "$SKIP_FULL" ] && exit 0          # launch list only
timeout 900 $PCMD > gpurun_out/plain2.log 2>&1 &&
timeout 1800 ncu --set full --clock-control none -k regex:tc_conv_slab -s ${NCU_SKIP:-82} -c ${NCU_COUNT:-13} -o gpurun_out/prof_chain $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
