#!/bin/bash
# Round-end style GPU pass: parity tests, smoke, bench, ncu launch list, one full ncu capture of the top kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 60 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 2>/dev/null | tee gpurun_out/bench_ref.json | cut -c1-300
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1ab|conv_tc" -s 10 -c 5 -o gpurun_out/prof_conv \
      python tools/bringup.py profile 480 752 64 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/
fi
