#!/usr/bin/env python
"""Minimal driver for profilers: N forwards (+ decode) of BUCTD-CoAM-W48 384x288 at batch B on cuda:0.
    ncu ... python tools/run_forward.py --iters 2 --batch 64"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_inputs  # noqa: E402
from buctd_b200 import inference, pose_hrnet, pose_hrnet_coam, synth  # noqa: E402
from buctd_b200.config import make_cfg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--variant", default="coam")
args = ap.parse_args()
cfg = make_cfg(args.variant, 48, (288, 384), cond="stacked" if args.variant == "coam" else "colored")
mod = pose_hrnet_coam if args.variant == "coam" else pose_hrnet
net = mod.get_pose_net(cfg, is_train=False, dtype=args.dtype)
net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)
net = net.cuda().eval()
dev = torch.device("cuda", 0)
if args.variant == "coam":
    rgb, kp, c, s = make_inputs(torch, cfg, args.batch, dev, 1234)
    with torch.no_grad():
        for _ in range(args.iters):
            preds, maxvals = net.predict_from_proposals(rgb, kp, c, s, post_process=True)     # the bench step
    torch.cuda.synchronize()
    print("done", tuple(preds.shape))
else:
    x = torch.randn(args.batch, 6, 384, 288, device=dev)
    with torch.no_grad():
        for _ in range(args.iters):
            hm = net(x)
            inference.decode_device(hm, None, None, post_process=True)
    torch.cuda.synchronize()
    print("done", tuple(hm.shape))
