python __graft_entry__.py --smoke 2>&1 | tail -2
for n in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_r04_${n}gpu.json 2> gpurun_out/bench_r04_${n}gpu.err
python -c "
import json; j=json.loads(open('gpurun_out/bench_r04_${n}gpu.json').read().strip().splitlines()[-1]); print(j['n_gpus'], round(j['value']), round(j['e2e']['value']), j['ms_per_step'], j['clocks'])"
done
