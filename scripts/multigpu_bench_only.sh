#!/bin/bash
# usage: scripts/multigpu_bench_only.sh N   (inside gpurun --gpus N): config 2 (default bench) and config 5 under torchrun
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_config2_n$N.json 2> gpurun_out/mg_config2_n$N.err
$TR bench.py --gpus $N --config 5 --steps 10 --warmup 3 > gpurun_out/mg_config5_n$N.json 2> gpurun_out/mg_config5_n$N.err
tail -c 300 gpurun_out/mg_config5_n$N.err gpurun_out/mg_config2_n$N.err
python - <<PY
import json
for c in (2, 5):
    try:
        d = json.loads(open(f"gpurun_out/mg_config{c}_n$N.json").read().strip().splitlines()[-1])
        print(c, {k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "e2e", "e2e_channel_mean", "allreduce", "clocks")})
    except Exception as e:
        print(c, "failed", e)
PY
