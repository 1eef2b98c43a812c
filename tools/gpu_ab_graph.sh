#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mnet.py -x -q -m gpu -k "network or engine or facade or pipeline or config or invarian or chain or sibling" > gpurun_out/graph_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/graph_tests.log
for g in 0 1 0 1; do
  VD_GRAPH=$g timeout 300 python bench.py --no-extras --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_graph$g.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_graph$g.json')); print('graph $g', d['value'], d['e2e']['value'], d['segments_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
VD_GRAPH=0 python tools/layer_times_mnet.py 16 1080 1920 5 bf16 2>/dev/null | grep "^forward"
VD_GRAPH=1 python tools/layer_times_mnet.py 16 1080 1920 5 bf16 2>/dev/null | grep "^forward"
