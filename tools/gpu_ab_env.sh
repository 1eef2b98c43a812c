#!/bin/bash
# A/B of an environment switch on the per-layer table:  tools/gpu_ab_env.sh VAR "v0 v1 ..." [B]
mkdir -p gpurun_out
VAR=$1; VALS=$2; B=${3:-16}
for v in $VALS; do
  env $VAR=$v timeout 300 python tools/layer_times.py $B 1080 1920 3 > gpurun_out/layers_${VAR}_$v.log 2>&1
  echo "== $VAR=$v"; head -1 gpurun_out/layers_${VAR}_$v.log; tail -2 gpurun_out/layers_${VAR}_$v.log | head -1
done
