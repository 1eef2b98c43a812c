#!/bin/bash
# first GPU contact: TMA probe, kernel parity tests, per-layer timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 "video-desensitization_b200/csrc/probe_tma" > gpurun_out/probe_tma.log 2>&1; echo "probe exit $?" >> gpurun_out/probe_tma.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/layer_times.py 8 1080 1920 3 > gpurun_out/layer_times.log 2>&1; echo "layer_times exit $?" >> gpurun_out/layer_times.log
tail -5 gpurun_out/probe_tma.log; tail -40 gpurun_out/pytest_gpu.log; tail -15 gpurun_out/layer_times.log
