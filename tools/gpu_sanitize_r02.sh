#!/bin/bash
# compute-sanitizer over the paths added in round 2: exact mode, descriptor-set matcher (tensor-core nomination + re-rank
# with re-scans), fused guided kernel with mapped-memory results, throughput outputs, the CUDA-graph replay of spfe_extract.
mkdir -p gpurun_out
T="tests/test_gpu_matcher.py tests/test_gpu_outputs.py tests/test_gpu_exact.py::test_exact_layers_match_oracle tests/test_guided.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -x -q -m gpu -k "not 4096 and not 2001" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_matcher.py tests/test_guided.py tests/test_gpu_exact.py::test_exact_layers_match_oracle -x -q -m gpu -k "near_tie or duplicates or sets or layers or deep" > gpurun_out/r02_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r02_racecheck.log
