timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-prof > gpurun_out/bench_latest.json 2>> gpurun_out/sweep.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_latest.json"))
print("value", round(d["value"]), "e2e", d["e2e"], "e2e_csr", d.get("e2e_csr"))
PY
