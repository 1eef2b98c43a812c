#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for f in 1 0; do echo "VD_STEM_POOL=$f"; VD_STEM_POOL=$f timeout 120 python tools/one_layer.py body.conv1 8 20 2>&1 | tail -1; done
timeout 300 python tools/layer_times.py 8 1080 1920 3 > gpurun_out/layer_times.log 2>&1; echo "layer_times exit $?" >> gpurun_out/layer_times.log
head -4 gpurun_out/layer_times.log; tail -4 gpurun_out/layer_times.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log | cut -c1-300
VD_STEM_POOL=0 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-200
