#!/bin/bash
# Final pass of the round: GPU suite, smoke, the default bench line, the exact-mode bench line (with its parity block),
# the reference arm and the ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench.json
timeout 600 python bench.py --exact > gpurun_out/r02_bench_exact.json 2> gpurun_out/r02_bench_exact.err; cut -c1-200 gpurun_out/r02_bench_exact.json
timeout 300 python bench.py --impl reference --steps 6 --warmup 3 2>/dev/null > gpurun_out/r02_bench_reference.json; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 300 sp_orb_slam_b200/lib/stream_bench tests/golden/superpoint_v1.spw 480 752 64 3 60 | tee gpurun_out/r02_native_stream_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0.05 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r02_launches.csv | head -30 | tee gpurun_out/r02_launch_shares.txt
