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
rgb, kp, c, s = make_inputs(torch, cfg, args.batch, dev, 1234)
with torch.no_grad():
    for _ in range(2):
        net.predict_from_proposals(rgb, kp, c, s, post_process=True)
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(args.reps):
        net.predict_from_proposals(rgb, kp, c, s, post_process=True)
    e1.record()
torch.cuda.synchronize()
fwd = e0.elapsed_time(e1) * 1e3 / args.reps
tot = sum(r[0] for r in rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    plan.run(stream)
e1.record()
torch.cuda.synchronize()
print("plan.run() alone: %.1f us (the recorded graph)" % (e0.elapsed_time(e1) * 1e3 / args.reps))
print("predict_from_proposals %.1f us; sum of plan ops (each alone, warm L2) %.1f us over %d ops" % (fwd, tot, len(rows)))


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
