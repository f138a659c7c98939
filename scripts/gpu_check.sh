#!/bin/bash
# Runs on the GPU box (via gpurun): per-file GPU parity tests, smoke(), a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in rules decoder encoder tc models; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -s > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$f.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 3000 gpurun_out/bench.log
if [ -n "${NCU_LIST}" ]; then
  PCMD="python bench.py --images 4 --profile-step"
  timeout 600 $PCMD > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?" | tee -a gpurun_out/summary.txt
fi
if [ -n "${NCU_FULL}" ]; then
  PCMD="python bench.py --images 4 --profile-step"
  timeout 600 $PCMD > gpurun_out/plain2.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-14} -o gpurun_out/prof_tc $PCMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a gpurun_out/summary.txt
fi
