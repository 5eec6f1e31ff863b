#!/usr/bin/env python
"""Times one chain launch (4 BasicBlocks = 8 convolutions 3x3 C->C of one HRNet branch, through a recorded plan) under the
regular plan and the dx-fused plans (BUCTD_DXF / BUCTD_DXF_MSUB are read when the op is added).
    python tools/chain_bench.py [--batch 64]"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from buctd_b200 import _lib as L  # noqa: E402
from gpu_utils import TDT, pack_w, to_act16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
L.load()
ctx = L.Context.get(0)
stream = torch.cuda.current_stream().cuda_stream
dt = L.F16
# variant = "<BUCTD_DXF>[:<BUCTD_DXF_MSUB>][/<BUCTD_CHAIN_DEBUG>][@<BUCTD_FORCE_MSUB>]"
CASES = [(48, 96, 72, ["0", "0@2", "0@1", "0/1", "1:1", "1:2", "1:2/1", "1:2/2", "1:2/4", "1:2/6", "1:2/7", "1:1/1", "1:1/2", "1:1/4", "1:1/7"]),
         (64, 96, 72, ["0", "0@1", "1:1", "1:1/1", "1:1/2", "1:1/4", "1:1/7"])]
EXTRA_KEYS = set()
if os.environ.get("CHAIN_BENCH_CASES"):
    CASES = eval(os.environ["CHAIN_BENCH_CASES"])
for (c, h, w, variants) in CASES:
    B = args.batch
    x = torch.randn(B, c, h, w, device="cuda")
    ws = [pack_w(torch.randn(c, c, 3, 3, device="cuda") / math.sqrt(9 * c), dt) for _ in range(8)]
    bs = [torch.randn(c, device="cuda").float().contiguous() * 0.1 for _ in range(8)]
    results = {}
    for var in variants:
        v0, *extra = var.split("+")              # "...+KEY=VAL": any other plan knob (BUCTD_CHAIN_PAIR, BUCTD_SW_MAX, ...)
        for key in [k for k in os.environ if k in EXTRA_KEYS]:
            del os.environ[key]
        for kv in extra:
            key, val = kv.split("=")
            EXTRA_KEYS.add(key)
            os.environ[key] = val
        for key, sep in (("BUCTD_CHAIN_LAG", "g"), ("BUCTD_FORCE_MSUB", "@"), ("BUCTD_CHAIN_DEBUG", "/"), ("BUCTD_DXF_MSUB", ":")):
            if sep in v0:
                v0, val = v0.split(sep)
                os.environ[key] = val
            else:
                os.environ.pop(key, None)
        os.environ["BUCTD_DXF"] = v0
        bufs = [to_act16(x, dt)] + [torch.zeros(B, h + 1, w + 1, c, dtype=TDT[dt], device="cuda") for _ in range(2)]
        a = L.TapGemmChainArgs()
        a.n_layers = 8
        cur = 0
        for blk in range(4):
            t, o = (cur + 1) % 3, (cur + 2) % 3
            for j, (src, dst, res) in enumerate(((cur, t, None), (t, o, cur))):
                g = a.layer[2 * blk + j]
                wp, n_alloc, k_cs = ws[2 * blk + j]
                g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, 9, 1, B
                g.out_h, g.out_w, g.in_h, g.in_w = h, w, h, w
                g.m_rows = B * (h + 1) * (w + 1)
                g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = c, c, c, n_alloc, k_cs
                g.out_cs, g.res_cs, g.relu, g.out_mode = c, (c if res is not None else 0), 1, L.OUT_ACT16
                g.a, g.w, g.bias, g.out = bufs[src].data_ptr(), wp.data_ptr(), bs[2 * blk + j].data_ptr(), bufs[dst].data_ptr()
                g.residual = bufs[res].data_ptr() if res is not None else None
            cur = o
        plan = L.Plan(ctx)
        plan.add(L.OP_CHAIN, a, "chain")
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
        flops = 8 * 2.0 * B * h * w * c * c * 9
        results[var] = bufs[cur].float().clone()
        d = (results[var] - results[variants[0]]).abs().max().item()
        print("C=%d %dx%d B=%d dxf=%-7s: %7.1f us per 8 convs (%5.1f us/conv), %6.1f TFLOP/s, max|diff vs regular| %.3g"
              % (c, h, w, B, var, us, us / 8, flops / us / 1e6, d), flush=True)
        del plan
