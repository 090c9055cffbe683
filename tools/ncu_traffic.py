#!/usr/bin/env python
"""Pair one step of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list with the kernel tags of a bench.py JSON line (same workload, same launch order) and write
profiles/dram_traffic.json + a markdown table.

    tools/ncu_traffic.py <launches.csv> <bench.json> <workload> <out_md>
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, bench_json, workload, out_md = sys.argv[1:5]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    d = per.setdefault(int(r[0]), {"name": r[4]})
    d[r[12]] = float(r[14]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1.0)
launches = [per[k] for k in sorted(per)]
bench = json.load(open(bench_json))
tags = list(bench["kernels"])
n = len(tags)
assert len(launches) >= n, (len(launches), n)
launches = launches[-n:]                      # the last full step of the capture
traffic, lines = {}, []
tot_ncu = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches)
tot_ev = sum(bench["kernels"][t]["ms_per_step"] for t in tags)
lines.append(f"| kernel tag | SASS kernel | event ms (bench) | share | ncu ms (cold, serialised) | share | algorithmic MB | DRAM MB read+written (ncu) |")
lines.append("|---|---|---|---|---|---|---|---|")
for t, l in zip(tags, launches):
    k = bench["kernels"][t]
    dram = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    traffic[t] = dram
    name = l["name"].split("(")[0].replace("void ", "")[:60]
    lines.append(f"| {t} | `{name}` | {k['ms_per_step']:.4f} | {k['ms_per_step']/tot_ev:.3f} | "
                 f"{l.get('gpu__time_duration.sum', 0):.4f} | {l.get('gpu__time_duration.sum', 0)/tot_ncu:.3f} | "
                 f"{k['bytes_per_launch']/1e6:.1f} | {dram/1e6:.1f} |")
path = os.path.join(ROOT, "profiles", "dram_traffic.json")
allt = json.load(open(path)) if os.path.exists(path) else {}
allt[workload] = traffic
json.dump(allt, open(path, "w"), indent=1)
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
