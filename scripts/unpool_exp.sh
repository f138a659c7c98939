#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tcx.py tests/test_gpu_encoder.py tests/test_gpu_resnet_tc.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_gradient.py -m gpu -x -q 2>&1 | tail -3
echo "bf16 $(LAYERS=2,4,1 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | tr '\n' '|')" | tee gpurun_out/unpool_exp.log
echo "fp32 $(PRECISION=fp32 LAYERS=2,4 REPS=9 timeout 200 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | sed 's/ (chunk 128)//' | tr '\n' '|')" | tee -a gpurun_out/unpool_exp.log
