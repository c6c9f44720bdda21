timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-gpu-eager --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('bench', round(j['value']), round(j['e2e']['value']), j['roofline']['avg_launch_us'], {k:v['ms'] for k,v in j['kernel_breakdown'].items() if k in ('mcab_decode_tc','dit_blocks')})"
python tools/bench_vae.py census 1024 2>&1 | tail -1 | cut -c1-400
