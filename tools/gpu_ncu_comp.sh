#!/bin/bash
mkdir -p gpurun_out
C=$1
python tools/composite_one.py $C > gpurun_out/comp_$C.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite_run -s 3 -c 1 -o gpurun_out/comp_$C python tools/composite_one.py $C > gpurun_out/ncu_comp_$C.log 2>&1
tail -2 gpurun_out/comp_$C.log
