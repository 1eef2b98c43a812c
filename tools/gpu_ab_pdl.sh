#!/bin/bash
mkdir -p gpurun_out
for g in 0 1 0 1; do
  VD_CONV_PDL=$g timeout 300 python bench.py --no-extras --steps 20 --warmup 3 2>gpurun_out/pdl_err$g.log | tail -1 > gpurun_out/bench_pdl$g.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$g.json')); print('pdl $g', d['value'], d['e2e']['value'], d['segments_ms']['network'], d['roofline']['frac'], d['clocks']['sm_mhz'])" || tail -3 gpurun_out/pdl_err$g.log
done
VD_CONV_PDL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "network or engine or chain or sibling or invarian" 2>&1 | tail -2
