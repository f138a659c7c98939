#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/b_ws.json 2> gpurun_out/b_ws.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/b_ws.json").readline())
print(round(d["value"]), d["ms_per_step"], d.get("breakdown_ms"), "e2e", round(d["e2e"]["value"]), d["clocks"], d["gpu_launches"])
PY
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-also 2> gpurun_out/b_ws3.err | cut -c1-300
