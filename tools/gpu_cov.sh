#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu -k "cov or batch_invariance or operator_call or shim" 2>&1 | tail -15 | tee gpurun_out/cov_tests.log
timeout 300 python tools/covstat.py 8 2>&1 | tail -8 | tee gpurun_out/covstat8.log
timeout 300 python tools/covstat.py 32 2>&1 | tail -8 | tee gpurun_out/covstat32.log
