#!/bin/bash
# A/B of composite_run_kernel occupancy variants (build/variants/libvdesens_minb{4,5}.so, -DVD_COMP_MINB) on the bench boxes
mkdir -p gpurun_out
L=video-desensitization_b200/vdesens_b200/libvdesens.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "nms or postprocess or detect or empty or facade" > gpurun_out/t_nms.log 2>&1; tail -2 gpurun_out/t_nms.log
cp $L /tmp/base.so
for v in base minb4 minb5 base; do
  if [ $v = base ]; then cp /tmp/base.so $L; else cp build/variants/libvdesens_$v.so $L; fi
  echo "== $v" >> gpurun_out/probe_ab.log
  timeout 300 python tools/composite_probe.py 2>&1 | grep -E "^all |^small|^big only  |^all mosaic" >> gpurun_out/probe_ab.log
done
cp /tmp/base.so $L
cat gpurun_out/probe_ab.log
timeout 600 python bench.py --steps 10 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-200
