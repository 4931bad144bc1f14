#!/usr/bin/env python
"""Aggregates `ncu --page source --print-source sass,cuda --csv` output per CUDA source line.
usage: ncu -i prof.ncu-rep --page source --print-source sass,cuda --csv | python tools/ncu_lines.py [N]"""
import csv
import sys

if sys.stdin.isatty() or (len(sys.argv) > 1 and sys.argv[1] in ("-h", "--help")):
    sys.exit(__doc__)
rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
data, cur, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        s = int(r[hdr.index("# Samples")])
        ins = int(r[hdr.index("Instructions Executed")])
        thr = int(r[hdr.index("Thread Instructions Executed")])
    except ValueError:
        continue
    data.append((s, ins, thr, cur, r[0], r[1][:96], r))
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
tott = sum(d[2] for d in data)
print(f"total samples {tot}  warp inst {toti}  thread inst {tott}  avg active lanes {tott / toti:.1f}")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for d in data:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(d[6][i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
for s, ins, thr, f, ln, text, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
    print(f"{s / tot * 100:5.1f}% inst {ins / toti * 100:5.1f}% lanes {thr / max(ins, 1):4.1f} {f}:{ln:>4} {text}  {[(b, a) for a, b in st]}")
