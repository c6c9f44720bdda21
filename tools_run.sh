# scratch script for gpurun sessions: GPU tests, then one bench line
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_latest.json 2>> gpurun_out/sweep.err; tail -c 400 gpurun_out/bench_latest.json
