python bench.py > gpurun_out/bench_r04_default.json 2> gpurun_out/bench_r04_default.err; python -c "
import json; j=json.load(open('gpurun_out/bench_r04_default.json')); print(round(j['value']), round(j['e2e']['value']), j['roofline']['frac'], j.get('gpu_eager_baseline'), j['cpu_baseline']['value'], j['clocks'])"
tail -3 gpurun_out/bench_r04_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r04_reference.json 2>/dev/null; cat gpurun_out/bench_r04_reference.json | cut -c1-200
