#!/bin/bash
# 8-GPU box: host-side copy ceiling (tools/d2h_probe.py) and the bench at 1/2/4/8 GPUs for the euroc and 1080p configs.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python tools/d2h_probe.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/d2h_probe.py 2>/dev/null; fi
done | tee gpurun_out/d2h_probe.txt
for cfg in euroc 1080p; do
for n in 8 4 2 1; do
  if [ $n = 1 ]; then timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-parity --no-cpu-baseline 2>gpurun_out/bench_${cfg}_n$n.err > gpurun_out/bench_${cfg}_n$n.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --config $cfg --gpus $n --steps 40 --warmup 5 2>gpurun_out/bench_${cfg}_n$n.err > gpurun_out/bench_${cfg}_n$n.json; fi
  python -c "
import json;d=json.load(open('gpurun_out/bench_${cfg}_n$n.json'));print('$cfg',$n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'full',round(d['e2e_full_outputs']['value']),'sus',round(d['sustained_value']),'d2h/step',d['e2e']['d2h_bytes_per_step'],d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done; done
