set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -3
python bench.py > gpurun_out/bench_r03_default.json 2> gpurun_out/bench_r03_default.err; tail -c 600 gpurun_out/bench_r03_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r03_reference.json 2>/dev/null
python tools/sweep.py gpurun_out/r03_sweep.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r03_ncu_launches.csv python bench.py --batch 1184 --chunk 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prof > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --cache-control none -k regex:dit_blocks --launch-skip 60 --launch-count 1 -o gpurun_out/r03_prof_dit_blocks -f python bench.py --batch 1184 --chunk 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prof > /dev/null 2>&1
ls -la gpurun_out | tail -8
