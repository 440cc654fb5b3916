#!/bin/bash
# First-contact diagnostics on the GPU box; every step is time-boxed and independent.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for spec in "layers 64 96 1" "layers 120 136 2" "layers 480 752 2" "match 801" "profile 480 752 16" "profile 480 752 32"; do
  tag=$(echo $spec | tr ' ' '_')
  echo "=== $spec" | tee gpurun_out/bringup_$tag.log
  timeout 420 python tools/bringup.py $spec >> gpurun_out/bringup_$tag.log 2>&1
  echo "exit $?" >> gpurun_out/bringup_$tag.log
  tail -n 45 gpurun_out/bringup_$tag.log
done
