#!/bin/bash
# Bring-up of the tensor-core conv1a path (conv1ab_mma.cuh): layer parity vs the oracle, A/B timing against the FFMA kernel.
mkdir -p gpurun_out
for mode in mma ffma; do
  echo "=== SPFE_CONV1=$mode layers" | tee gpurun_out/c1m_layers_$mode.log
  SPFE_CONV1=$mode timeout 300 python tools/bringup.py layers 120 136 2 2>&1 | grep -v "^  \(conv[234]\|heat\|semi\|dense\)" | head -40 | tee -a gpurun_out/c1m_layers_$mode.log
  echo "=== SPFE_CONV1=$mode profile" | tee gpurun_out/c1m_profile_$mode.log
  SPFE_CONV1=$mode timeout 300 python tools/bringup.py profile 480 752 32 2>&1 | tail -22 | tee -a gpurun_out/c1m_profile_$mode.log
done
