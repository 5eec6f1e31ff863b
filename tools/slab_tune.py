#!/usr/bin/env python
"""Sweep the slab tap-GEMM plan space (N tile, super-tile size) for the W48 BasicBlock shapes:
    python tools/slab_tune.py        (uses the BUCTD_FORCE_BN / BUCTD_FORCE_MSUB test knobs per process)"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.environ.get("SLAB_TUNE_CASE")
if CASE is None:
    shapes = {"48,48,96,72": [48], "96,96,48,36": [96], "192,192,24,18": [192, 64, 128], "384,384,12,9": [192, 128, 64, 256],
              "64,64,96,72": [64], "64,256,96,72,1": [256, 128, 64], "256,64,96,72,1": [64]}
    for shape, bns in shapes.items():
        for bn in bns:
            for msub in (1, 2, 3, 4):
                env = dict(os.environ, SLAB_TUNE_CASE=shape, BUCTD_FORCE_BN=str(bn), BUCTD_FORCE_MSUB=str(msub), BUCTD_PLAN_DEBUG="1")
                r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
                plan = [l for l in r.stderr.splitlines() if l.startswith("[buctd] slab plan")]
                print(shape, "bn", bn, "msub<=", msub, "|", r.stdout.strip(), "|", plan[0].split("->")[1].strip() if plan else r.stderr[-300:])
    sys.exit(0)

sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import math  # noqa: E402

import torch  # noqa: E402

from buctd_b200 import _lib as L  # noqa: E402
from gpu_utils import TDT, pack_w, to_act16  # noqa: E402

v = [int(t) for t in CASE.split(",")]
cin, cout, h, w = v[:4]
k = v[4] if len(v) > 4 else 3
B, dt = 64, L.F16
ctx = L.Context.get(0)
stream = torch.cuda.current_stream().cuda_stream
x = torch.randn(B, cin, h, w, device="cuda")
wt = torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(k * k * cin)
bias = torch.randn(cout, device="cuda").float().contiguous()
a16 = to_act16(x, dt)
out = torch.zeros(B, h + 1, w + 1, cout, dtype=TDT[dt], device="cuda")
r16 = to_act16(torch.randn(B, cout, h, w, device="cuda"), dt)
wp, n_alloc, k_cs = pack_w(wt, dt)
g = L.TapGemmArgs()
g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, k * k, 1, B
g.out_h, g.out_w, g.in_h, g.in_w = h, w, h, w
g.m_rows = B * (h + 1) * (w + 1)
g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = cin, a16.shape[-1], cout, n_alloc, k_cs
g.out_cs, g.res_cs, g.relu, g.out_mode = out.shape[-1], r16.shape[-1], 1, L.OUT_ACT16
g.a, g.w, g.bias, g.out, g.residual = a16.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), r16.data_ptr()
plan = L.Plan(ctx)
plan.add(L.OP_TAPGEMM, g, "conv")
for _ in range(3):
    plan.run(stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    plan.run(stream)
e1.record()
torch.cuda.synchronize()
print("%.1f us" % (e0.elapsed_time(e1) * 1e3 / 20))
