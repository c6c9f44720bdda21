mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 -p no:cacheprovider 2>&1 | tail -4
for c in 392 784 1024; do
  python bench.py --steps 2 --warmup 2 --batch 2352 --chunk $c --no-cpu-baseline --no-e2e > gpurun_out/bench_c$c.json 2>> gpurun_out/sweep.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_c$c.json"))
print("chunk", $c, "value", round(d["value"]), "model_tflops", d["model_tflops"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"])
print("   ", {k:(v["share"], round(v["ms"]/v["launches"]*1000,1)) for k,v in list(d["kernel_breakdown"].items())[:9]})
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 40 -c 8 -o gpurun_out/prof_gemm_v2 python bench.py --steps 1 --warmup 1 --batch 392 --chunk 392 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/ncu_v2.log 2>&1; tail -2 gpurun_out/ncu_v2.log; ls -la gpurun_out/*.ncu-rep
