#!/bin/bash
# Round-2 profiling pass: ncu launch list of the bench command, then one `ncu --set full` capture of every kernel.
# gpurun copies back at most 64 MiB: the raw metric page is exported to CSV on the box and an oversized report is dropped.
mkdir -p gpurun_out
if [ "${1:-}" != "full-only" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0.05 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
fi
timeout 1500 ncu --set full --clock-control none -c 120 -o gpurun_out/r02_prof_all \
    python tools/r02_kernels.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/r02_prof_all.ncu-rep --page raw --csv > gpurun_out/r02_prof_all_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r02_prof_all.ncu-rep)
if [ "$sz" -gt 45000000 ]; then rm gpurun_out/r02_prof_all.ncu-rep; echo "report $sz bytes: dropped, raw CSV kept"; fi
ls -la gpurun_out/ | head -30
