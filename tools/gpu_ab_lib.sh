#!/bin/bash
# A/B of library variants on the per-layer table:  tools/gpu_ab_lib.sh "base variantname base variantname" [B]
#   variants live in build/variants/libvdesens_<name>.so (built by hand with an extra -D), "base" = the in-tree library
mkdir -p gpurun_out
L=video-desensitization_b200/vdesens_b200/libvdesens.so
cp $L /tmp/base.so
B=${2:-16}; i=0
for v in $1; do
  i=$((i+1))
  if [ $v = base ]; then cp /tmp/base.so $L; else cp build/variants/libvdesens_$v.so $L; fi
  timeout 300 python tools/layer_times.py $B 1080 1920 3 > gpurun_out/layers_lib_${v}_$i.log 2>&1
  echo "== $v"; head -1 gpurun_out/layers_lib_${v}_$i.log; tail -2 gpurun_out/layers_lib_${v}_$i.log | head -1
done
cp /tmp/base.so $L
