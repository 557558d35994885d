#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counts by CUDA source line (first matching launch).
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top-n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, tot, fpath, seen_funcs = {}, 0, "", 0
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iN = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= iN: continue
    if r[0] == "":            # SASS line
        continue
    try: n = int(r[iN]); s = int(r[iS] or 0)
    except ValueError: continue
    key = (fpath, r[0], r[1].strip())
    a = agg.setdefault(key, [0, 0]); a[0] += n; a[1] += s
# several launches repeat the same lines: the counts add up; report shares
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-instructions (all captured launches of the kernel) {tot}, samples {ts}")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*n/tot:5.1f}% inst {100*s/max(ts,1):5.1f}% smp  {key[0]}:{key[1]:>4s}: {key[2][:120]}")
