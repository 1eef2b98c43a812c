#!/bin/bash
# ingest tests, HBM read/write-mix probe, and the bench's ingest segment at 1 / 2 / 4 rows per CTA
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ingest" > gpurun_out/ingest_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/ingest_tests.log
python tools/probes/hbm_mix.py 2>&1 | tee gpurun_out/r02_hbm_mix_probe.log
for r in 1 2 4 1 2 4; do
  VD_INGEST_ROWS=$r timeout 300 python bench.py --no-extras --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_ingest_rows$r.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_ingest_rows$r.json')); print('rows $r', d['segments_ms']['ingest_with_frame_copy'], d['stage_hbm']['ingest']['frac_of_hbm_peak'], d['value'])"
done
