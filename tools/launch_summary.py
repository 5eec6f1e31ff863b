#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list.
    python tools/launch_summary.py list.csv [--skip N] [--take N] [--group grid] [--filter name]"""
import argparse, collections, csv, re
ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--take", type=int, default=0)
ap.add_argument("--group", default="")
ap.add_argument("--filter", default="")
ap.add_argument("--top", type=int, default=45)
args = ap.parse_args()
rows = list(csv.reader(open(args.csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, mi, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    e = d.setdefault(int(r[0]), {"k": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("buctd::", ""), "g": r[gi]})
    v = float(r[vi].replace(",", ""))
    if r[mi].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    elif r[ui] in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    e[r[mi]] = v
ids = list(d)
ids = ids[args.skip:]
if args.take:
    ids = ids[:args.take]
t = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i in ids:
    e = d[i]
    if args.filter and args.filter not in e["k"]:
        continue
    key = e["k"] + ((" grid " + e["g"]) if args.group == "grid" else "")
    t[key][0] += 1
    t[key][1] += e.get("gpu__time_duration.sum", 0.0)
    t[key][2] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in t.values())
print("%d launches, %.3f ms" % (sum(v[0] for v in t.values()), tot / 1e3))
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print("%9.1f us %5d %5.1f%% %8.1f us/launch %8.1f MB/launch %6.2f TB/s  %s" % (v[1], v[0], 100 * v[1] / tot, v[1] / v[0], v[2] / v[0] / 1e6,
                                                                                 v[2] / max(v[1], 1e-9) / 1e6, k[:100]))
