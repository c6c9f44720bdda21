python tools/kernel_timeline.py 1184 2>&1 | grep -A1 "^mlp1" | grep "\["
