for c in 128 152 203 304 608; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --chunk $c > gpurun_out/bench_chunk_$c.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/bench_chunk_$c.log"):
    if l.startswith("{"):
        d=json.loads(l); print($c, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["breakdown_ms"], round(sum(x["ms"] for x in d["roofline"]["layers"])/$c*128,3))
PY
done
