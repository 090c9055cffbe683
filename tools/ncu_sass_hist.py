#!/usr/bin/env python
"""Opcode histogram (dynamic, warp-level instructions executed) from `ncu --page source --csv`."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else None
kern = None; hdr = None; agg = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; agg[kern] = collections.Counter(); hdr = None; continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr and kern and len(r) > 5:
        src = r[hdr["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else src[:10]
        try: n = int(r[hdr["Instructions Executed"]])
        except: n = 0
        agg[kern][op] += n
for k, c in agg.items():
    if want and want not in k: continue
    tot = sum(c.values())
    print("==", k[:120], "total warp-inst", tot)
    for op, n in c.most_common(28):
        print(f"   {op:10s} {n:12d} {100.0*n/tot:5.1f}%")
