"""Summarise one kernel of an .ncu-rep (read here, no GPU) into a small JSON for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep <kernel regex> "<description>" "<command>" > profiles/rNN_ncu_x_summary.json"""
import csv
import io
import json
import re
import subprocess
import sys

rep, pat, desc, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum|sm__pipe_tensor.*cycles_active|"
                  r"sm__inst_executed_pipe_tensor|sm__warps_active\.avg\.pct|launch__(registers_per_thread|block_size|grid_size|shared_mem_per_block_dynamic)|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|smsp__issue_active\.avg\.pct|sm__throughput\.avg\.pct|lts__t_bytes\.sum$|"
                  r"smsp__inst_executed\.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|sm__cycles_elapsed\.max|lts__t_sectors_op_(read|write)\.sum$")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name")
sel = [r for r in rows[2:] if re.search(pat, r[name_col])]
res = {"kernel": desc, "command": cmd, "launches_in_report": len(sel), "metrics": {}}
if sel:
    r = sel[-1]
    for i, h in enumerate(hdr):
        if KEEP.search(h):
            res["metrics"][h] = {"value": r[i], "unit": units[i]}
print(json.dumps(res, indent=1))
