import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# split by "Kernel Name" records
inst = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; inst.append(cur); continue
    if r and r[0] == "Address": hdr = r; continue
    if cur is not None and r: cur.append(r)
print("instances", len(inst), [len(x) for x in inst])
k = inst[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iW = hdr.index("L1 Wavefronts Shared"); iWI = hdr.index("L1 Wavefronts Shared Ideal")
tot = sum(int(r[iI]) for r in k); ts = sum(int(r[iS]) for r in k); tw = sum(int(r[iW]) for r in k); twi = sum(int(r[iWI]) for r in k)
print("total inst", tot, "samples", ts, "smem wavefronts", tw, "ideal", twi)
seg_i = seg_s = seg_w = seg_wi = 0; n = 0; start = 0
ops = {}
for j, r in enumerate(k):
    op = r[1].split()
    op = [o for o in op if not o.startswith('@')][0] if op else ''
    seg_i += int(r[iI]); seg_s += int(r[iS]); seg_w += int(r[iW]); seg_wi += int(r[iWI]); n += 1
    ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + int(r[iI])
    if op.startswith("BAR") or op.startswith("EXIT") or j == len(k) - 1:
        if seg_i > tot * 0.002 or True:
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
            print(f"sass {start:4d}-{j:4d} ({n:4d} instrs): inst {100*seg_i/tot:5.1f}%  smp {100*seg_s/ts:5.1f}%  wavefronts {100*seg_w/max(tw,1):5.1f}% (ideal {100*seg_wi/max(tw,1):5.1f}%)  {r[1].strip()[:40]}  " + " ".join(f"{a}:{100*b/tot:.1f}" for a, b in top))
        seg_i = seg_s = seg_w = seg_wi = 0; n = 0; start = j + 1; ops = {}
