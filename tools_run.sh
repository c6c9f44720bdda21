python bench.py > gpurun_out/bench_r04_default.json 2> gpurun_out/bench_r04_default.err; python -c "
import json; j=json.load(open('gpurun_out/bench_r04_default.json')); print(round(j['value']), round(j['e2e']['value']), j['roofline']['frac'], j['gpu_eager_baseline']['value'], j['cpu_baseline']['value'], j['clocks'], j['gpu_launches'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r04_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r04_ncu_launches.csv python bench.py --batch 1184 --chunk 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prof --no-gpu-eager > /dev/null 2>&1
wc -l gpurun_out/r04_ncu_launches.csv
