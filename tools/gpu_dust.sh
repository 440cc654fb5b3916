#!/bin/bash
# one short GPU call for the dust-pose path: parity tests, the C++ shim, latency, memcheck
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_pose_dust.py -m gpu -q -s -p no:cacheprovider > gpurun_out/dust_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/dust_pytest.log
tail -25 gpurun_out/dust_pytest.log
timeout 60 python tools/dust_pose_time.py 100 > gpurun_out/dust_time.json 2> gpurun_out/dust_time.err; cat gpurun_out/dust_time.json; tail -3 gpurun_out/dust_time.err
timeout 90 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/dust_pose_time.py 2 > gpurun_out/dust_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/dust_memcheck.log
