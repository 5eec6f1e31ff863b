"""A few training steps of the bench configuration (for ncu / timing experiments).
    python tools/run_train.py --steps 2 --batch 32 [--eager]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buctd_b200 import condition, pose_hrnet_coam, synth
from buctd_b200.config import make_cfg
from buctd_b200.train import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--eager", action="store_true")
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()
cfg = make_cfg("coam", 48, (288, 384), cond="stacked")
net = pose_hrnet_coam.get_pose_net(cfg, is_train=False, dtype=args.dtype)
net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)
net = net.cuda().train()
B = args.batch
ts = TrainStep(net, B, dtype=args.dtype, drop_p=0.1, graph=not args.eager)
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.empty(B, 20, 384, 288, device="cuda")
x[:, :3] = torch.randn(B, 3, 384, 288, generator=g, device="cuda")
kp = torch.rand(B, 17, 2, generator=g, device="cuda") * 250 + 8
condition.rasterize_condition(kp, 384, 288, "stacked", out=x, c_off=3)
tgt = torch.rand(B, 17, 96, 72, generator=g, device="cuda") ** 6
tw = torch.ones(B, 17, 1, device="cuda")
torch.cuda.synchronize()
for i in range(args.steps):
    t0 = time.perf_counter()
    loss = ts.step(x, tgt, tw)
    torch.cuda.synchronize()
    print("step %d loss %.6f %.1f ms" % (i, float(loss.item()), (time.perf_counter() - t0) * 1e3), flush=True)
