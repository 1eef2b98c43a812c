#!/bin/bash
# usage: gpu_ncu_layer.sh <layer> <tag> [env]
mkdir -p gpurun_out
L=$1; T=$2
python tools/one_layer.py $L 8 3 > gpurun_out/one_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 1 -o gpurun_out/layer_$T python tools/one_layer.py $L 8 3 > gpurun_out/ncu_$T.log 2>&1
cat gpurun_out/one_$T.log | tail -1
