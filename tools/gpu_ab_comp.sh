#!/bin/bash
# A/B of library variants on the compositing probe:  tools/gpu_ab_comp.sh "base mb3 base mb5"
mkdir -p gpurun_out
L=video-desensitization_b200/vdesens_b200/libvdesens.so
cp $L /tmp/base.so
i=0
for v in $1; do
  i=$((i+1))
  if [ $v = base ]; then cp /tmp/base.so $L; else cp build/variants/libvdesens_$v.so $L; fi
  timeout 300 python tools/composite_probe.py > gpurun_out/comp_lib_${v}_$i.log 2>&1
  echo "== $v"; grep -E "^all |^big only  |^small \(fused\)|^all mosaic" gpurun_out/comp_lib_${v}_$i.log
done
cp /tmp/base.so $L
