timeout 600 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|vae_|many cells|Error|error|assert" | head -30
python bench.py --steps 2 --warmup 2 --batch 2352 --chunk 784 --no-cpu-baseline --no-e2e > gpurun_out/bench_v6.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_v6.json"))
print("value", round(d["value"]), "model_tflops", d["model_tflops"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
print("   ", {k:(v["share"], round(v["ms"]/v["launches"]*1000,1)) for k,v in list(d["kernel_breakdown"].items())[:10]})
PY
