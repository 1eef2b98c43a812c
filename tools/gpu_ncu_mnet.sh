#!/bin/bash
# usage: gpu_ncu_mnet.sh <op name> <tag>   — plain run first, then ncu --set full on the LAST of the 18 mnet_* launches (14 in the forward pass + 4 isolated)
mkdir -p gpurun_out
OP=$1; T=$2
python tools/one_mnet_op.py $OP 16 3 > gpurun_out/one_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mnet_ -s 17 -c 1 -o gpurun_out/mnet_$T python tools/one_mnet_op.py $OP 16 3 > gpurun_out/ncu_$T.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/one_$T.log
