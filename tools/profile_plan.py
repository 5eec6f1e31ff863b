#!/usr/bin/env python
"""Per-op timing of one recorded forward plan with CUDA events (each op replayed alone, back to back).
    python tools/profile_plan.py --batch 64 [--reps 5] [--top 40]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_inputs  # noqa: E402
from buctd_b200 import inference, pose_hrnet, pose_hrnet_coam, synth  # noqa: E402
from buctd_b200.config import make_cfg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--dtype", default="fp16")
args = ap.parse_args()
cfg = make_cfg("coam", 48, (288, 384), cond="stacked")
net = pose_hrnet_coam.get_pose_net(cfg, is_train=False, dtype=args.dtype)
net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)
net = net.cuda().eval()
dev = torch.device("cuda", 0)
x, c, s = make_inputs(torch, cfg, args.batch, dev, 1234)
with torch.no_grad():
    for _ in range(2):
        hm = net(x)
torch.cuda.synchronize()
st = net._plans[(args.batch, 0)]
plan = st["plan"]
stream = torch.cuda.current_stream().cuda_stream
rows = []
for i, label in enumerate(plan.labels):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.run(stream, i, i + 1)
    e0.record()
    for _ in range(args.reps):
        plan.run(stream, i, i + 1)
    e1.record()
    torch.cuda.synchronize()
    rows.append((e0.elapsed_time(e1) * 1e3 / args.reps, i, label))
# whole forward
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(args.reps):
        hm = net(x)
    e1.record()
torch.cuda.synchronize()
fwd = e0.elapsed_time(e1) * 1e3 / args.reps
tot = sum(r[0] for r in rows)
# plan only (one call, all ops back to back) and decode
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    plan.run(stream)
e1.record()
torch.cuda.synchronize()
print("plan.run() alone: %.1f us (ops back to back in one call)" % (e0.elapsed_time(e1) * 1e3 / args.reps))
e0.record()
for _ in range(args.reps):
    inference.decode_device(hm, c, s, post_process=True)
e1.record()
torch.cuda.synchronize()
print("decode: %.1f us" % (e0.elapsed_time(e1) * 1e3 / args.reps))
# head ops of forward (not in the plan): time a forward that stops right after the first plan op
from buctd_b200 import _lib as L  # noqa: E402
ctx = st["ctx"]
H, W, Cx = 384, 288, x.shape[1]
f32 = net._packed["f32"]
def timeit(fn, name):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%s: %.1f us" % (name, e0.elapsed_time(e1) * 1e3 / args.reps))
s1 = st["s1"]
sa = L.StemArgs()
sa.dtype, sa.batch, sa.x_channels, sa.h, sa.w, sa.out_cs = net._dt, args.batch, Cx, H, W, s1.cs
sa.x, sa.weight, sa.bias, sa.out = x.data_ptr(), f32["conv1.w"].data_ptr(), f32["conv1.b"].data_ptr(), s1.ptr
timeit(lambda: ctx.run(L.OP_STEM, sa, stream), "stem conv1")
for i, cond in st["cond"].items():
    ra = L.ResizeArgs()
    ra.dtype, ra.batch, ra.x_channels, ra.c_off, ra.cc = net._dt, args.batch, Cx, 3, cond.c
    ra.h, ra.w, ra.oh, ra.ow, ra.out_cs, ra.antialias = H, W, cond.h, cond.w, cond.cs, 1
    ra.x, ra.out = x.data_ptr(), cond.ptr
    timeit(lambda ra=ra: ctx.run(L.OP_RESIZE, ra, stream), "resize -> %dx%d" % (cond.h, cond.w))
fs = st["final_src"]
out = torch.empty(args.batch, 17, fs.h, fs.w, device=dev)
fa = net._tapgemm("final_layer", fs, out.data_ptr(), 0, fs.h, fs.w, out_mode=L.OUT_NCHW32)
timeit(lambda: ctx.run(L.OP_TAPGEMM, fa, stream), "final layer (immediate tapgemm incl. tensor-map encode)")
print("forward %.1f us; sum of plan ops (alone, warm L2) %.1f us over %d ops" % (fwd, tot, len(rows)))


def klass(label):
    import re
    m = re.match(r"(stage\d)\.\d+\.branches\.(\d)\.\d\.(conv\d)", label)
    if m:
        return "%s branch%s %s" % (m.group(1)[:5] + "#", m.group(2), m.group(3))
    return re.sub(r"\d+", "#", label)


agg = collections.defaultdict(lambda: [0.0, 0])
for us, i, label in rows:
    a = agg[klass(label)]
    a[0] += us
    a[1] += 1
print("--- by class")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
    print("%9.1f us  %5.1f%%  n=%3d  avg %7.1f  %s" % (us, 100 * us / tot, n, us / n, k))
print("--- top single ops")
for us, i, label in sorted(rows, reverse=True)[:25]:
    print("%9.1f us  #%d %s" % (us, i, label))
