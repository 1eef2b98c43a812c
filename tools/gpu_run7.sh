#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "fused" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fused.log
tail -25 gpurun_out/pytest_fused.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/layer_times.py 8 1080 1920 3 > gpurun_out/layer_times.log 2>&1; echo "layer_times exit $?" >> gpurun_out/layer_times.log
head -4 gpurun_out/layer_times.log; tail -4 gpurun_out/layer_times.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log | cut -c1-300
