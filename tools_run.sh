for e in 6 14; do
  echo "EXP=$e"; SCLDM_EXP=$e python bench.py --no-cpu-baseline --no-e2e --no-gpu-eager --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['value']), j['roofline']['avg_launch_us'], j['roofline']['frac'])"
done
SCLDM_EXP=14 python tools/kernel_timeline.py 1184 2>&1 | grep -A1 "median cycles" | grep "\["
