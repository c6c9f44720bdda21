timeout 300 python -m pytest tests/test_gpu_dit.py -m gpu -q -x --timeout=120 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/bench_mega1.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_mega1.json"))
print("single value", round(d["value"]), "ms", d["ms_per_step"])
PY
SCLDM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_dit.py -m gpu -q -x --timeout=120 -p no:cacheprovider -k "large_batch or sample_ode or batch_inv" 2>&1 | tail -2
SCLDM_PAIR=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/bench_pair.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_pair.json"))
print("pair value", round(d["value"]), "ms", d["ms_per_step"])
PY
