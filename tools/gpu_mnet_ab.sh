#!/bin/bash
# MobileNet variant: kernel tests, then per-op times at the depthwise run lengths
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mnet.py -x -q -m gpu > gpurun_out/mnet_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/mnet_tests.log
for ny in ${1:-4 2 1}; do
  VD_MNET_DW_NY=$ny timeout 300 python tools/layer_times_mnet.py 16 1080 1920 5 bf16 > gpurun_out/mnet_layers_ny$ny.log 2>&1
  echo "== NY=$ny"; grep -E "^mobilenet|^stem|^dwpw|^dw |^forward" gpurun_out/mnet_layers_ny$ny.log
done
timeout 300 python tools/layer_times_mnet.py 16 1080 1920 5 fp16 > gpurun_out/mnet_layers_fp16.log 2>&1
grep -E "^mobilenet|^forward" gpurun_out/mnet_layers_fp16.log
