#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:grid_attn -c 12 --csv --log-file gpurun_out/attn_launches.csv python scripts/dec_probe.py > gpurun_out/attn_ncu.log 2>&1
grep "grid_attn" gpurun_out/attn_launches.csv | awk -F'","' '{print $5, $NF}' | tail -6
