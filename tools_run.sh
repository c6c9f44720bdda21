mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01_default.json 2> gpurun_out/bench_r01_default.err; tail -c 1500 gpurun_out/bench_r01_default.json; tail -2 gpurun_out/bench_r01_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_r01_default.err; cat gpurun_out/bench_r01_reference.json | head -c 600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 2200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 1 --batch 784 --chunk 784 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log; wc -l gpurun_out/r01_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_|attn16|mcab_decode_tc|final_step" -s 60 -c 14 -o gpurun_out/r01_prof_top python bench.py --steps 1 --warmup 1 --batch 784 --chunk 784 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
