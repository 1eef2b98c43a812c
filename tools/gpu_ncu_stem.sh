#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "fused" 2>&1 | tail -3
python tools/one_layer.py body.conv1 8 3 > gpurun_out/one_sp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_pool -s 1 -c 1 -o gpurun_out/layer_sp python tools/one_layer.py body.conv1 8 3 > gpurun_out/ncu_sp.log 2>&1
tail -1 gpurun_out/one_sp.log
