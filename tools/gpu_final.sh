#!/bin/bash
# last GPU pass of the round: the whole GPU suite, smoke, the default bench line, dust-pose timing + ncu capture
mkdir -p gpurun_out
timeout 120 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 60 python tools/dust_pose_time.py 100 > gpurun_out/dust_time.json 2> gpurun_out/dust_time.err; cat gpurun_out/dust_time.json; tail -3 gpurun_out/dust_time.err
timeout 60 ncu --set full --clock-control none --import-source on -k regex:dust_pose -c 1 -o gpurun_out/prof_dust2 python tools/dust_pose_time.py 2 > gpurun_out/ncu_dust2.log 2>&1; tail -2 gpurun_out/ncu_dust2.log | cut -c1-200
timeout 150 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
