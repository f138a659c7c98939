#!/bin/bash
mkdir -p gpurun_out
for sl in 200 1000 200 1000 50; do
  LRPX_TC_SLEEP=$sl timeout 300 python bench.py --config 2 --steps 15 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('sleep=$sl', round(d['value']), round(d['ms_per_step'],2), d['breakdown_ms']['encoder_relevance_chain'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/sleep_step.log
