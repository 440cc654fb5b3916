mkdir -p gpurun_out
for cfg in euroc; do
for n in 8 4 2 1; do
  if [ $n = 1 ]; then timeout 300 python bench.py --config $cfg --steps 40 --warmup 5 --no-parity --no-cpu-baseline 2>gpurun_out/bench_${cfg}_n$n.err > gpurun_out/bench_${cfg}_n$n.json
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --config $cfg --gpus $n --steps 40 --warmup 5 2>gpurun_out/bench_${cfg}_n$n.err > gpurun_out/bench_${cfg}_n$n.json; fi
  python -c "
import json;d=json.load(open('gpurun_out/bench_${cfg}_n$n.json'));print('$cfg',$n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'full',round(d['e2e_full_outputs']['value']),'sus',round(d['sustained_value']),'roof',round(d['roofline']['frac'],3),d['clocks']['sm_mhz'])"
done; done
