timeout 900 python -m pytest tests/test_gpu_vae.py -m gpu -q -x --timeout=600 -p no:cacheprovider -s -k "latent_token_order or full_size" 2>&1 | grep -E "latent|passed|failed|Error|assert" | tail -8
for m in heun2 dopri5; do
python bench.py --method $m --no-cpu-baseline --no-gpu-eager --steps 2 --warmup 3 > gpurun_out/bench_r04_$m.json 2>gpurun_out/bench_r04_$m.err; python -c "
import json; j=json.load(open('gpurun_out/bench_r04_$m.json')); print('$m', round(j['value']), round(j['e2e']['value']), j['config']['dit_evaluations_per_solve'], j['roofline']['frac'], j['model_tflops'])" || tail -3 gpurun_out/bench_r04_$m.err
done
