#!/bin/bash
mkdir -p gpurun_out
echo "== shared queues" | tee gpurun_out/e2e_probe.log
timeout 400 python tools/e2e_probe.py 2>&1 | tail -8 | tee -a gpurun_out/e2e_probe.log
echo "== SPFE_SLOT_STREAMS=1" | tee -a gpurun_out/e2e_probe.log
SPFE_SLOT_STREAMS=1 timeout 400 python tools/e2e_probe.py 2>&1 | tail -8 | tee -a gpurun_out/e2e_probe.log
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
