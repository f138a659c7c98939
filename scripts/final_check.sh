#!/bin/bash
# Round-end check on one B200: the whole GPU suite, then the driver-style default bench (both arms).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_all.log
tail -4 gpurun_out/gpu_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").readline())
print(round(d["value"]), d["ms_per_step"], d.get("breakdown_ms"), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["clocks"])
for k, v in d.get("also", {}).items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("unavailable", ""))
print([ (l["layer"], l["ms"]) for l in d["roofline"]["layers"]])
PY
