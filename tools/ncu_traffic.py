#!/usr/bin/env python
"""Summarises an `ncu --set full` capture of the chain tap-GEMM launches (ncu -i X.ncu-rep --page raw --csv) into
profiles/dominant_traffic.json: per launch the kernel name, device time, DRAM bytes read + written, tensor-pipe and
L2 hit rates.  bench.py takes `roofline.traffic` of the dominant kernel from that file.
    ncu -i gpurun_out/r02_chain_full.ncu-rep --page raw --csv > raw.csv; python tools/ncu_traffic.py raw.csv out.json"""
import csv
import json
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "%": 1.0}
rows = list(csv.reader(open(sys.argv[1])))
head, unit, body = rows[0], rows[1], rows[2:]


def col(r, name):
    i = head.index(name)
    return float(r[i].replace(",", "")) * SCALE.get(unit[i], 1.0)


out = []
for r in body:
    out.append({"kernel": r[head.index("Kernel Name")].split("(")[0].replace("void ", ""),
                "grid": r[head.index("Grid Size")], "block": r[head.index("Block Size")],
                "us": round(col(r, "gpu__time_duration.sum"), 2),
                "dram_read_bytes": col(r, "dram__bytes_read.sum"), "dram_write_bytes": col(r, "dram__bytes_write.sum"),
                "tensor_pipe_pct": round(col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 2),
                "l2_hit_pct": round(col(r, "lts__t_sector_hit_rate.pct"), 2)})
json.dump({"source": "ncu --set full --clock-control none, tools/run_forward.py --iters 2 --batch 64 (cold L2, isolated launches)",
           "launches": out}, open(sys.argv[2], "w"), indent=1)
for o in out:
    print(o)
