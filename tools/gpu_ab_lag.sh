#!/bin/bash
# composite tests on the new schedule, then the probe on the library variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "composite or config5 or dirty or streaming_pipeline or gauss" > gpurun_out/lag_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/lag_tests.log
bash tools/gpu_ab_comp.sh "${1:-base head}"
