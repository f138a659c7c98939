#!/bin/bash
# quick GPU iteration: selected tests + bench variants given in $BENCHES (semicolon separated arg strings)
mkdir -p gpurun_out
for f in ${TESTS}; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -s -x > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?"; tail -n 4 gpurun_out/test_$f.log
done
i=0
IFS=';' read -ra BL <<< "${BENCHES}"
for b in "${BL[@]}"; do
  timeout 900 python bench.py $b > gpurun_out/bench_$i.log 2>&1
  echo "bench[$b] exit $?"; tail -c 2500 gpurun_out/bench_$i.log; echo
  i=$((i+1))
done
