#!/bin/bash
# Round-2 A/B on one B200: first-layer row walk (LRPX_TC_WALK) and the re-laid-out decoder attention rule.
# Parity tests first (bounded by timeout: a barrier bug in a persistent kernel would otherwise hang the box).
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_tc.py tests/test_gpu_decoder.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/walk_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/walk_tests.log
tail -5 gpurun_out/walk_tests.log
grep -q "rc=0" gpurun_out/walk_tests.log || exit 1
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/b_walk.json 2> gpurun_out/b_walk.err
LRPX_TC_WALK=0 timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/b_nowalk.json 2> gpurun_out/b_nowalk.err
python - <<'PY'
import json
for f in ("b_walk", "b_nowalk"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").readline())
        print(f, round(d["value"]), d["ms_per_step"], d.get("breakdown_ms"), [l for l in d["roofline"]["layers"] if l["layer"] == 0])
    except Exception as e:
        print(f, "failed", e)
PY
