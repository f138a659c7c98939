for v in 1 0 1 0; do
  echo "== LRPX_TC_INPUT3=$v"
  LRPX_TC_INPUT3=$v LAYERS=0 REPS=9 python scripts/one_layer.py 2>&1 | grep layer
done
for v in 1 0; do
  LRPX_TC_INPUT3=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('INPUT3=$v', round(d['value']), d['ms_per_step'], d['breakdown_ms']['encoder_relevance_chain'])"
done
