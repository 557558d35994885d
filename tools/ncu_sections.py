#!/usr/bin/env python
"""Per-section totals (instructions, shared-memory wavefronts, stall samples) of one kernel in an ncu report.
usage: tools/ncu_sections.py <report> <kernel-regex> <file> name:from-to [name:from-to ...]"""
import csv, io, subprocess, sys
rep, kre, fname = sys.argv[1:4]
sections = {}
for a in sys.argv[4:]:
    n, rng = a.split(":"); lo, hi = rng.split("-"); sections[n] = (int(lo), int(hi))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cols = ["Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "# Samples"]
fpath = ""; hdr = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = [hdr.index(c) for c in cols]; continue
    if hdr is None or r[0] == "" or len(r) <= max(ix): continue
    try: vals = [int(r[i] or 0) for i in ix]
    except ValueError: continue
    key = (fpath, int(r[0]))
    acc = per.setdefault(key, [0] * len(cols))
    for k, v in enumerate(vals): acc[k] += v
tot = [sum(v[k] for v in per.values()) for k in range(len(cols))]
print(f"{'section':10s} {'inst%':>6s} {'warp-inst':>12s} {'thr/inst':>8s} {'smem wf':>12s} {'ideal':>12s} {'samples%':>8s}")
def show(name, acc):
    print(f"{name:10s} {100*acc[0]/max(tot[0],1):6.1f} {acc[0]:12d} {acc[1]/max(acc[0],1):8.1f} {acc[2]:12d} {acc[3]:12d} {100*acc[4]/max(tot[4],1):8.1f}")
other = [0] * len(cols)
accs = {n: [0] * len(cols) for n in sections}
for (f, l), v in per.items():
    hit = False
    if f == fname:
        for n, (lo, hi) in sections.items():
            if lo <= l <= hi:
                for k in range(len(cols)): accs[n][k] += v[k]
                hit = True; break
    if not hit:
        for k in range(len(cols)): other[k] += v[k]
for n in sections: show(n, accs[n])
show("other", other); show("TOTAL", tot)
