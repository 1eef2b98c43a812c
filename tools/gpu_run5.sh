#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for b in 1 2 4; do
  timeout 300 python tools/layer_times.py $b 1080 1920 3 > gpurun_out/layer_times_B$b.log 2>&1; echo "exit $?" >> gpurun_out/layer_times_B$b.log
  head -1 gpurun_out/layer_times_B$b.log; tail -3 gpurun_out/layer_times_B$b.log | head -1
done
