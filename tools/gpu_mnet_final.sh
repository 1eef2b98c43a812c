#!/bin/bash
# MobileNet variant, artefacts of record: kernel tests, per-op times (bf16 + fp16), FMA issue-rate probe, ncu captures of
# two fused blocks, a depthwise layer and the stem
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mnet.py -x -q -m gpu > gpurun_out/mnet_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/mnet_tests.log
python tools/layer_times_mnet.py 16 1080 1920 10 bf16 2>/dev/null | grep -v Warning > gpurun_out/r02_layer_times_mnet_B16_1080p.log
python tools/layer_times_mnet.py 16 1080 1920 10 fp16 2>/dev/null | grep -v Warning > gpurun_out/r02_layer_times_mnet_B16_1080p_fp16.log
head -1 gpurun_out/r02_layer_times_mnet_B16_1080p.log; grep "^forward" gpurun_out/r02_layer_times_mnet_B16_1080p.log
head -1 gpurun_out/r02_layer_times_mnet_B16_1080p_fp16.log; grep "^forward" gpurun_out/r02_layer_times_mnet_B16_1080p_fp16.log
./build/fma_rate > gpurun_out/r02_fma_rate_probe.log 2>&1
bash tools/gpu_ncu_mnet.sh body.stage1.3 r02dwpw3c
bash tools/gpu_ncu_mnet.sh body.stage1.1 r02dwpw1c
bash tools/gpu_ncu_mnet.sh body.stage1.5.dw r02dw5c
bash tools/gpu_ncu_mnet.sh body.stage1.0 r02stemc
