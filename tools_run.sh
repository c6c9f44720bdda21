timeout 900 python -m pytest tests/test_gpu_dit.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/bench_mega1.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_mega1.json"))
print("mega value", round(d["value"]), "ms", d["ms_per_step"])
PY
python tools/kernel_timeline.py 1184 2>&1 | grep -A1 "^mlp1\|^qkv" | grep "\["
