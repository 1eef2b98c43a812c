#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for f in 1 0; do echo "PATCH_WRES=$f"; for l in body.layer1.1.conv2 ssh1.conv5X5_2 ssh2.conv7X7_2; do VD_CONV_PATCH_WRES=$f timeout 120 python tools/one_layer.py $l 8 20 2>&1 | tail -1; done; done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log | cut -c1-200
VD_CONV_PATCH_WRES=0 timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-200
