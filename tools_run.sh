timeout 900 python -m pytest tests/test_gpu_dit.py tests/test_gpu_vae.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_v13.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_v13.json"))
print("value", round(d["value"]), "ms", d["ms_per_step"], "model_tflops", d["model_tflops"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
print("   ", {k:(v["share"], round(v["ms"]/v["launches"]*1000,1)) for k,v in list(d["kernel_breakdown"].items())[:10]})
PY
