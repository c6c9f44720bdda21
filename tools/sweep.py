"""Batch sweep (BASELINE configs[1]) and dataset-shape sweep (configs[2], [3]) of bench.py on one GPU.
Usage (GPU box): python tools/sweep.py [out.jsonl]     -- one bench.py JSON line per configuration."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sweep.jsonl")
runs = [("dentate_gyrus", b) for b in (64, 256, 1024, 4096, 16384)] + [(d, 2368) for d in ("hlca", "tabula_muris", "replogle", "parse1m")]
with open(out, "w") as f:
    for dataset, batch in runs:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--dataset", dataset, "--batch", str(batch), "--steps", "3", "--warmup", "3",
               "--no-cpu-baseline", "--no-gpu-eager"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            j = json.loads(line)
        except ValueError:
            print(dataset, batch, "FAILED", r.stderr[-2000:])
            continue
        f.write(json.dumps(j) + "\n")
        f.flush()
        print(f"{dataset:14s} B={batch:6d} value {j['value']:9.0f} e2e {j['e2e']['value']:9.0f} e2e_csr {j.get('e2e_csr', {}).get('value', 0):9.0f} "
              f"ode {j['stages']['ode_only_cells_per_s']:9.0f} decode {j['stages']['decode_only_cells_per_s']:10.0f} "
              f"frac {j['roofline']['frac']:.3f} top {j['roofline']['kernel']}")
