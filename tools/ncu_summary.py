#!/usr/bin/env python
"""Turn an .ncu-rep into the small CSV kept under profiles/: one row per metric, one column per launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_kernel.csv [regex ...]

Keeps launch geometry, duration, DRAM/L2/shared traffic, issue and pipe utilisation, occupancy limits
and the stall breakdown (metric names matching the default list or the regexes given)."""
import csv
import io
import re
import subprocess
import sys

KEEP = [r"^Kernel Name$", r"^Block Size$", r"^Grid Size$", r"gpu__time_duration", r"^dram__", r"^lts__t_(bytes|sector)", r"lts__throughput",
        r"^launch__", r"smsp__issue_active", r"smsp__inst_executed\.sum$", r"sm__inst_executed\.sum$", r"sm__throughput", r"sm__warps_active",
        r"sm__pipe_(fma|alu|fmaheavy|fmalite)_cycles_active", r"issue_stalled.*per_issue_active", r"l1tex__data_pipe_lsu_wavefronts_mem_shared",
        r"sm__cycles_elapsed", r"gpc__cycles_elapsed.max", r"smsp__thread_inst_executed_per_inst", r"l1tex__t_bytes", r"sm__sass_inst_executed_op_shared"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pats = [re.compile(p) for p in KEEP + sys.argv[3:]]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        for i, name in enumerate(hdr):
            if any(p.search(name) for p in pats):
                w.writerow([name, units[i]] + [d[i] for d in data])


if __name__ == "__main__":
    main()
