#!/bin/bash
mkdir -p gpurun_out
for mode in 1 0; do
  echo "=== SPFE_PAIR=$mode layers" | tee gpurun_out/pair_layers_$mode.log
  SPFE_PAIR=$mode timeout 200 python tools/bringup.py layers 120 136 2 2>&1 | grep -E "conv1b|conv2a|conv2b|conv3a|keypoint sets|Error|error|timeout" | head -12 | tee -a gpurun_out/pair_layers_$mode.log
  echo "=== SPFE_PAIR=$mode profile" | tee gpurun_out/pair_profile_$mode.log
  SPFE_PAIR=$mode timeout 200 python tools/bringup.py profile 480 752 64 2>&1 | tail -14 | tee -a gpurun_out/pair_profile_$mode.log
done
