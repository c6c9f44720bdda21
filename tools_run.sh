timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -4
for occ in 2 3; do
  echo "DEC_OCC=$occ"; SCLDM_DEC_OCC=$occ python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['e2e']['value']), j['roofline']['avg_launch_us'], {k:v['ms'] for k,v in j['kernel_breakdown'].items() if k in ('final_step_tc','nb_finalize','mcab_decode_tc','dec_latent')})"
done
