#!/bin/bash
# A/B of VD_BRANCH (SSH levels 2/3 on side streams): parity subset with the switch on, then the interleaved forward A/B
mkdir -p gpurun_out
VD_BRANCH=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "network or engine or facade or pipeline or invarian or sibling" > gpurun_out/branch_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/branch_tests.log
timeout 300 python tools/ab_forward.py VD_BRANCH 0 1 16 9 2>/dev/null | tail -1 | tee gpurun_out/branch_ab.log
