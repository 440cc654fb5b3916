#!/bin/bash
# First GPU call of the next round: what round 1 left unverified.
#   1. racecheck of dust_pose_kernel alone (round 1 ran out of GPU budget before it completed)
#   2. the whole GPU suite + smoke + default bench line
mkdir -p gpurun_out
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/dust_pose_min.py > gpurun_out/dust_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/dust_racecheck.log
timeout 180 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
