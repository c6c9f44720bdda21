set -x
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err; tail -c 600 gpurun_out/bench_r02_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 400 gpurun_out/bench_r02_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --batch 1184 --chunk 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/ncu_b1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:dit_blocks --launch-skip 60 --launch-count 1 -o gpurun_out/r02_prof_dit_blocks -f python bench.py --batch 1184 --chunk 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-prof > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out | tail -8
