#!/bin/bash
mkdir -p gpurun_out
for kps in 0 2 4 7; do for occ in 2 1; do
  echo "KPS=$kps OCC=$occ"; VD_CONV_KPS=$kps VD_CONV_OCC=$occ timeout 120 python tools/one_layer.py body.conv1 8 20 2>&1 | tail -1
done; done > gpurun_out/stem.log 2>&1
cat gpurun_out/stem.log
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "conv or stem or network or ingest" 2>&1 | tail -3
