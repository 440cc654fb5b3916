mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python tools/d2h_probe.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/d2h_probe.py 2>/dev/null; fi
done | tee gpurun_out/d2h_probe.txt
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 60 --warmup 5 2>gpurun_out/bench_n$n.err > gpurun_out/bench_n$n.json
  python -c "
import json;d=json.load(open('gpurun_out/bench_n$n.json'));print($n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'full',round(d['e2e_full_outputs']['value']),'sus',round(d['sustained_value']),d['clocks'])"
done
timeout 200 python -m pytest tests/test_gpu_outputs.py -x -q -m gpu 2>&1 | tail -3
