#!/usr/bin/env python
"""Micro-benchmark of lanes: one BasicBlock step (conv1 + conv2 with residual) of the four W48 branches,
serial on one lane with all SMs vs. concurrent on four lanes with SM shares, both as instantiated graphs.
    python tools/lane_bench.py [--shares 52,36,32,28] [--depth 4]"""
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
ap.add_argument("--shares", default="52,36,32,28")
ap.add_argument("--depth", type=int, default=4, help="BasicBlocks per branch")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--branches", type=int, default=4)
args = ap.parse_args()
ctx = L.Context.get(0)
stream = torch.cuda.current_stream().cuda_stream
B = args.batch
GEO = [(48, 96, 72), (96, 48, 36), (192, 24, 18), (384, 12, 9)][:args.branches]
dt = L.F16
keep = []


def make_branch(c, h, w):
    bufs = [to_act16(torch.randn(B, c, h, w, device="cuda"), dt) for _ in range(3)]
    wt = torch.randn(c, c, 3, 3, device="cuda") / math.sqrt(9 * c)
    wp, n_alloc, k_cs = pack_w(wt, dt)
    bias = torch.randn(c, device="cuda").float().contiguous()
    keep.append((bufs, wp, bias))

    def op(src, dst, res):
        g = L.TapGemmArgs()
        g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, 9, 1, B
        g.out_h, g.out_w, g.in_h, g.in_w = h, w, h, w
        g.m_rows = B * (h + 1) * (w + 1)
        g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = c, bufs[0].shape[-1], c, n_alloc, k_cs
        g.out_cs, g.res_cs, g.relu, g.out_mode = bufs[0].shape[-1], (bufs[0].shape[-1] if res is not None else 0), 1, L.OUT_ACT16
        g.a, g.w, g.bias, g.out = bufs[src].data_ptr(), wp.data_ptr(), bias.data_ptr(), bufs[dst].data_ptr()
        g.residual = bufs[res].data_ptr() if res is not None else None
        return g
    ops = []
    cur = 0
    for _ in range(args.depth):
        t, o = (cur + 1) % 3, (cur + 2) % 3
        ops.append(op(cur, t, None))
        ops.append(op(t, o, cur))
        cur = o
    return ops


branches = [make_branch(*g) for g in GEO]
shares = [int(v) for v in args.shares.split(",")]


def build(lanes):
    plan = L.Plan(ctx)
    for i, ops in enumerate(branches):
        for g in ops:
            plan.add(L.OP_TAPGEMM, g, "b%d" % i, lane=i if lanes else 0, max_ctas=shares[i] if lanes else 0)
    plan.instantiate()
    return plan


def timeit(plan):
    for _ in range(3):
        plan.run(stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        plan.run(stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / args.reps


serial = build(False)
t_serial = timeit(serial)
conc = build(True)
t_conc = timeit(conc)
print("depth %d x %d branches, B=%d: serial %.1f us, lanes %s %.1f us (%.2fx)" % (args.depth, len(GEO), B, t_serial, shares[:len(GEO)], t_conc,
                                                                                 t_serial / t_conc))
# each branch alone at its share (what the lanes would cost if they did not interfere)
for i, ops in enumerate(branches):
    plan = L.Plan(ctx)
    for g in ops:
        plan.add(L.OP_TAPGEMM, g, "b%d" % i, lane=0, max_ctas=shares[i])
    plan.instantiate()
    print("   branch %d alone with %d CTAs: %.1f us" % (i, shares[i], timeit(plan)))
