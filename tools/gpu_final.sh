#!/bin/bash
# end-of-round evidence: full GPU suite, per-layer table, default bench line (copied into profiles/ by hand)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/layer_times.py 8 1080 1920 3 > gpurun_out/layer_times.log 2>&1; echo "layer_times exit $?" >> gpurun_out/layer_times.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_final.log
tail -2 gpurun_out/bench_final.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
