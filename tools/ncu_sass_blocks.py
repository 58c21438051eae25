#!/usr/bin/env python3
"""Basic-block execution profile from `ncu -i rep --page source --csv --print-source sass --kernel-id :::N`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = his[which]; end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
print("kernel:", rows[hi - 1][1][:80] if hi > 0 else "?", f"({which + 1} of {len(his)})")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) > idx["Thread Instructions Executed"] and r[0].startswith("0x")]
I = lambda r, k: int(r[idx[k]] or 0)
tot_inst = sum(I(r, "Instructions Executed") for r in data); tot_thr = sum(I(r, "Thread Instructions Executed") for r in data)
print("total warp inst", tot_inst, "thread inst", tot_thr, "avg lanes %.2f" % (tot_thr / tot_inst), "n sass", len(data))
blocks = []
for i, r in enumerate(data):
    ie, te, sm = I(r, "Instructions Executed"), I(r, "Thread Instructions Executed"), I(r, "# Samples")
    op = r[idx["Source"]].strip().split()
    op = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else ""))
    if blocks and abs(blocks[-1]["ie"] - ie) <= 0.002 * max(ie, 1):
        b = blocks[-1]; b["n"] += 1; b["samples"] += sm; b["end"] = i; b["ops"].append(op); b["te"] += te
    else:
        blocks.append(dict(start=i, end=i, ie=ie, te=te, n=1, samples=sm, ops=[op]))
ts = sum(b["samples"] for b in blocks) or 1
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
for b in blocks:
    share = b["ie"] * b["n"] / tot_inst
    if share > minshare:
        lanes = b["te"] / max(b["ie"] * b["n"], 1)
        print(f"[{b['start']:4d}-{b['end']:4d}] n={b['n']:3d} exec={b['ie']:10d} lanes={lanes:5.1f} inst-share={share:5.3f} stall-samples={b['samples']/ts:5.3f}  {' '.join(b['ops'][:16])}")
