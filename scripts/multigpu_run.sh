#!/bin/bash
# usage: scripts/multigpu_run.sh N   (inside gpurun --gpus N): probe + config 5 + default config-2 bench under torchrun
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
$TR scripts/multigpu_probe.py > gpurun_out/mg_probe_n$N.json 2> gpurun_out/mg_probe_n$N.err
$TR bench.py --gpus $N --config 5 --steps 10 --warmup 3 > gpurun_out/mg_config5_n$N.json 2> gpurun_out/mg_config5_n$N.err
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_config2_n$N.json 2> gpurun_out/mg_config2_n$N.err
tail -c 300 gpurun_out/mg_probe_n$N.err gpurun_out/mg_config5_n$N.err gpurun_out/mg_config2_n$N.err
wc -c gpurun_out/mg_*_n$N.json
