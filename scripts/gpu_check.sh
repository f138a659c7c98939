#!/bin/bash
# Runs on the GPU box (via gpurun): per-file GPU parity tests, smoke(), a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${TESTS:-rules decoder encoder tc models beam ablation fullsize}; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -s > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" | tee -a gpurun_out/summary.txt
  grep -v "mbarrier wait timed out" gpurun_out/test_$f.log | tail -n 3
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench exit $?" | tee -a gpurun_out/summary.txt
tail -c 3500 gpurun_out/bench.log
