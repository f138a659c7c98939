for v in 0 1 0 1; do
  LRPX_TC_CLUSTER=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cl_$v.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/bench_cl_$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("cluster=$v", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["breakdown_ms"])
PY
done
LRPX_TC_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -2
