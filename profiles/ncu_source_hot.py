#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report
(--import-source on):  python profiles/ncu_source_hot.py rep.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys
from collections import defaultdict


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main(path, pat="", top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    per_line = defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, text
    fname, func, hdr, take, done = "", "", None, False, set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
            take = pat in func
        elif r[0] == "Line No":
            hdr = r
        elif take and hdr and r[0].isdigit():
            key = (fname, int(r[0]))
            if (func, key) in done:
                continue
            i_inst, i_thr, i_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            e = per_line[key]
            e[0] += num(r[i_inst])
            e[1] += num(r[i_thr])
            e[2] += num(r[i_smp])
            e[3] = r[1].strip()[:90]
    tot_i = sum(e[0] for e in per_line.values()) or 1
    tot_s = sum(e[2] for e in per_line.values()) or 1
    print(f"# {path} kernel~'{pat}': {tot_i} warp instructions, {tot_s} stall samples (all captured launches)")
    print("# file:line  inst%  avg_threads  samples%  source")
    for key, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: int(top)]:
        print(f"{key[0]}:{key[1]:<5d} {100*e[0]/tot_i:5.1f}%  {e[1]/max(e[0],1):5.1f}  {100*e[2]/tot_s:5.1f}%  {e[3]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
