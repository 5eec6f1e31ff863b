#!/usr/bin/env python
"""Micro-benchmark of single tap-GEMM launches through a recorded plan (no host prepare in the timed
region): time per launch back to back, against the tensor floor (N/2 cycles per 128xNx16 MMA) and,
with a -DBUCTD_ROLE_TIMERS build (BUCTD_LIB=build/libbuctd_b200_timers.so), where each role waits.
    python tools/conv_sweep.py [--quick]"""
import argparse
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from buctd_b200 import _lib as L  # noqa: E402
from gpu_utils import TDT, pack_w, to_act16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--nk", action="store_true", help="N / K separation experiment: 64->48, 48->64, 32->32, 64->32, 32->64 at 96x72")
args = ap.parse_args()
lib = L.load()
ctx = L.Context.get(0)
stream = torch.cuda.current_stream().cuda_stream
# cin, cout, k, B, h, w, residual
CASES = [(48, 48, 3, 64, 96, 72, False), (48, 48, 3, 64, 96, 72, True), (96, 96, 3, 64, 48, 36, True),
         (192, 192, 3, 64, 24, 18, True), (384, 384, 3, 64, 12, 9, True), (64, 256, 1, 64, 96, 72, True),
         (256, 64, 1, 64, 96, 72, False), (64, 64, 3, 64, 96, 72, False)]
if args.nk:
    CASES = [(48, 48, 3, 64, 96, 72, False), (64, 48, 3, 64, 96, 72, False), (48, 64, 3, 64, 96, 72, False), (64, 64, 3, 64, 96, 72, False),
             (32, 32, 3, 64, 96, 72, False), (64, 32, 3, 64, 96, 72, False), (32, 64, 3, 64, 96, 72, False)]
elif not args.quick:
    CASES += [(48, 48, 3, b, 96, 72, False) for b in (4, 8, 16, 32, 128)]
    CASES += [(96, 96, 3, b, 48, 36, False) for b in (8, 32, 128)]
    CASES += [(192, 192, 3, b, 24, 18, False) for b in (8, 32, 128)]
keep = []
for (cin, cout, k, B, h, w, res) in CASES:
    dt = L.F16
    x = torch.randn(B, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(k * k * cin)
    bias = torch.randn(cout, device="cuda").float().contiguous()
    a16 = to_act16(x, dt)
    out = torch.zeros(B, h + 1, w + 1, cout, dtype=TDT[dt], device="cuda")
    r16 = to_act16(torch.randn(B, cout, h, w, device="cuda"), dt) if res else None
    wp, n_alloc, k_cs = pack_w(wt, dt)
    g = L.TapGemmArgs()
    g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, k * k, 1, B
    g.out_h, g.out_w, g.in_h, g.in_w = h, w, h, w
    g.m_rows = B * (h + 1) * (w + 1)
    g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = cin, a16.shape[-1], cout, n_alloc, k_cs
    g.out_cs, g.res_cs, g.relu, g.out_mode = out.shape[-1], (r16.shape[-1] if res else 0), 1, L.OUT_ACT16
    g.a, g.w, g.bias, g.out = a16.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr()
    g.residual = r16.data_ptr() if res else None
    plan = L.Plan(ctx)
    plan.add(L.OP_TAPGEMM, g, "conv")
    keep.append((a16, out, r16, wp, bias, plan))
    for _ in range(3):
        plan.run(stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        plan.run(stream)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.reps
    rows = B * (h + 1) * (w + 1)
    tiles = -(-rows // 128)
    bn = cout if cout <= 256 else 192
    nt = -(-cout // bn)
    nmma_total = tiles * k * k * (-(-cin // 16)) * nt
    floor_us = nmma_total * (bn / 2.0) / 148.0 / 1965.0
    flops = 2.0 * B * h * w * cin * cout * k * k
    print("%3d->%3d k%d B%3d %2dx%2d res=%d: %7.1f us/launch  %6.1f TFLOP/s  tensor-floor %5.1f us (%.0f MMAs/SM, %.0f cyc/MMA achieved)"
          % (cin, cout, k, B, h, w, res, us, flops / us / 1e6, floor_us, nmma_total / 148.0, us * 1965.0 / (nmma_total / 148.0)))
    if hasattr(lib, "buctd_debug_role_cycles"):
        plan.run(stream)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (148 * 16))()
        lib.buctd_debug_role_cycles(buf, 148 * 16)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 16)
        m, mx = a.mean(0), a.max(0)
        print("      MMA thread: total %.0f (max %.0f) = wait acc_empty %.0f + wait slab %.0f + wait W %.0f + issue %.0f | epilogue: total %.0f, "
              "wait acc_full %.0f, wait stage %.0f, compute %.0f | CTA: prologue %.0f, after pdl_wait %.0f, whole %.0f (max %.0f), kernel %.0f"
              % (m[0], mx[0], m[1], m[2], m[3], m[6], m[4], m[5], m[9], m[10], m[11], m[12], m[14], mx[14], us * 1965.0))
