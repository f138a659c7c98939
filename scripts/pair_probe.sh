#!/bin/bash
# cta_group::2 pair mode: correctness under the tc test files, then the bench layer table with the mode off / on / on for all widths
for P in 1 2; do
  LRPX_TC_PAIR=$P timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcx.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
done
for P in 0 1 2; do
  LRPX_TC_PAIR=$P timeout 600 python bench.py --no-also --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_pair$P.json 2> gpurun_out/bench_pair$P.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pair$P.json").read().strip().splitlines()[-1])
print("PAIR=$P value", round(d["value"]), "ms", round(d["ms_per_step"],2), d["breakdown_ms"], "frac", round(d["roofline"]["frac"],3))
print("   layers ms:", [(l["layer"], l["ms"], l["tflops"]) for l in d["roofline"].get("layers", [])])
PY
done
