"""Where does the gradient error of the 16-bit training step come from?  (loss scale, batch / map size, library bf16)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from buctd_b200 import pose_hrnet, pose_hrnet_coam, synth
from buctd_b200.arch import HRNetArch
from buctd_b200.config import make_cfg
from buctd_b200.train import TrainStep
from oracle import train_oracle as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def data(B, w, h, seed=4321):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, 3, h, w, generator=g, device="cuda")
    tgt = torch.rand(B, 17, h // 4, w // 4, generator=g, device="cuda") ** 4
    tw = (torch.rand(B, 17, 1, generator=g, device="cuda") < 0.8).float()
    return x, tgt, tw


def err(grads, ref, keys=None):
    ks = keys or list(ref)
    num = sum(float((grads[k].double() - ref[k].double()).norm()) ** 2 for k in ks)
    den = sum(float(ref[k].double().norm()) ** 2 for k in ks)
    return np.sqrt(num / den)


EARLY = ["conv1.weight", "conv2.weight", "layer1.0.conv1.weight", "layer1.1.conv1.weight", "layer1.2.conv1.weight"]
LATE = ["final_layer.weight", "stage4.2.branches.0.3.conv2.weight", "stage4.2.branches.0.3.conv1.weight"]

for (B, w, h) in ((4, 96, 128), (8, 192, 256)):
    cfg = make_cfg("plain", 32, (w, h))
    sd = synth.synth_state_dict(HRNetArch(cfg), 0)
    x, tgt, tw = data(B, w, h)
    sdc = {k: v.cuda() for k, v in sd.items()}
    _, _, ref, _ = T.train_forward_backward(sdc, cfg, x, tgt, tw)
    ref = {k: v.cpu() for k, v in ref.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, _, lib, _ = T.train_forward_backward(sdc, cfg, x, tgt, tw)
    lib = {k: v.float().cpu() for k, v in lib.items()}
    print("B %d %dx%d  library bf16 autocast: global %.4f early %.4f late %.4f" % (B, w, h, err(lib, ref), err(lib, ref, EARLY), err(lib, ref, LATE)))
    for dtype, ls in (("fp16", 128.0), ("fp16", 1024.0), ("fp16", 65536.0), ("bf16", 1.0)):
        net = pose_hrnet.get_pose_net(cfg, is_train=False, dtype=dtype)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train()
        ts = TrainStep(net, B, dtype=dtype, loss_scale=ls, drop_p=0.0, graph=False)
        ts.forward_backward(x, tgt, tw)
        torch.cuda.synchronize()
        g = {k: p.grad.detach().float().cpu() for k, p in net.named_parameters()}
        print("   %s scale %7.0f: global %.4f early %.4f late %.4f   conv1 %.4f" % (dtype, ls, err(g, ref), err(g, ref, EARLY), err(g, ref, LATE), err(g, ref, ["conv1.weight"])))
