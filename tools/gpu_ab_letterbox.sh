#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "ingest or letterbox or facade or mixed or upscale" > gpurun_out/ingest_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/ingest_tests.log
for r in 1 0 1 0; do echo "== staged rows: $r"; VD_INGEST_LETTERBOX_ROWS=$r python tools/letterbox_time.py 2>/dev/null | tail -3; done
