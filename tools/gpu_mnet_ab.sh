#!/bin/bash
# MobileNet variant: kernel tests, then per-op times for each environment setting given ("A=1,B=2" per argument; "-" = defaults)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mnet.py -x -q -m gpu > gpurun_out/mnet_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/mnet_tests.log
i=0
for cfg in "${@:--}"; do
  i=$((i+1))
  envs=""; [ "$cfg" != "-" ] && envs=$(echo $cfg | tr ',' ' ')
  env $envs timeout 300 python tools/layer_times_mnet.py 16 1080 1920 5 bf16 > gpurun_out/mnet_layers_$i.log 2>&1
  echo "== $cfg"; grep -E "^mobilenet|^stem|^dwpw|^dw |^forward" gpurun_out/mnet_layers_$i.log
done
