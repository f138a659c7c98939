#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the default bench workload (one warm-up + one step), (2) --set full on
# the relevance-chain kernels of one chunk (13 launches) of that same command.
mkdir -p gpurun_out
PCMD="python bench.py --profile-step --no-graph"
timeout 900 $PCMD > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 $PCMD > gpurun_out/plain2.log 2>&1 &&
timeout 1800 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s ${NCU_SKIP:-155} -c ${NCU_COUNT:-13} -o gpurun_out/prof_chain $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
