#!/bin/bash
for dbg in 0 128 0 128; do
  echo "== LRPX_TC_DEBUG=$dbg (128: no L2 prefetch of gain rows)"
  LRPX_TC_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['breakdown_ms']['encoder_relevance_chain'], [(x['layer'],x['ms']) for x in d['roofline']['layers']])"
done
