#!/usr/bin/env python
"""Times the channel attention's fc_o (nn.Linear over the planar [B*C, H*W] matrix, `buctd_tapgemm` ROWS16) on the three
CoAM branches of W48 at 384x288 under the CTA-pair plan (pairgemm.cu) and the one-CTA slab plan (BUCTD_PAIRGEMM=0).
    python tools/pairgemm_bench.py [--batch 64]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from buctd_b200 import _lib as L  # noqa: E402
from gpu_utils import TDT, pack_w, rup  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
L.load()
ctx = L.Context.get(0)
dt = L.F16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (c, h, w) in [(48, 96, 72), (96, 48, 36), (192, 24, 18)]:
    rows, k, n = args.batch * c, (h + 1) * (w + 1), h * w
    lda, ldo = rup(k, 8), rup(n, 8)
    a16 = (torch.randn(rows, lda, device="cuda")).to(TDT[dt])
    wp, n_alloc, k_cs = pack_w((torch.randn(n, k, 1, 1, device="cuda") / math.sqrt(k)), dt)
    bias = torch.randn(n, device="cuda")
    out = torch.zeros(rows, ldo, dtype=TDT[dt], device="cuda")
    res = {}
    for pair in ("1", "0"):
        os.environ["BUCTD_PAIRGEMM"] = pair
        g = L.TapGemmArgs()
        g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, 1, 1, 1
        g.m_rows, g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = rows, k, lda, n, n_alloc, k_cs
        g.out_cs, g.relu, g.out_mode = ldo, 0, L.OUT_ROWS16
        g.a, g.w, g.bias, g.out = a16.data_ptr(), wp.data_ptr(), bias.data_ptr(), out.data_ptr()
        plan = L.Plan(ctx)
        plan.add(L.OP_TAPGEMM, g, "fc_o")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            plan.run(st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(st)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res[pair] = (ts[len(ts) // 2], out.float().abs().sum().item())
    fl = 2.0 * rows * k * n
    print("C=%3d HW=%4d  M=%5d K=%4d N=%4d  pair %7.1f us (%6.1f TFLOP/s)   one-CTA %7.1f us (%6.1f TFLOP/s)   checksum %s" % (
        c, n, rows, k, n, res["1"][0], fl / res["1"][0] * 1e-6, res["0"][0], fl / res["0"][0] * 1e-6,
        "equal" if abs(res["1"][1] - res["0"][1]) <= 1e-3 * abs(res["0"][1]) else "DIFFERENT %r %r" % (res["1"][1], res["0"][1])))
