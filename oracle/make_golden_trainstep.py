"""TEST INFRASTRUCTURE -- writes tests/golden/train_step_*.npz from the UNMODIFIED reference in train() mode (row a-T):
model.train() -> forward -> JointsMSELoss(use_target_weight=True) -> backward (lib/core/function.py:113-138) on seeded
synthetic crops / weights, CPU fp32.  nn.Dropout modules are put in eval() (their mask comes from torch's RNG, which the
CUDA path cannot share; the dropout arithmetic is checked separately against the oracle with a shared hash mask).
Stored: loss, heat-maps, per-parameter gradient norms, the first 12 entries of every gradient, and the running statistics
after the step.        python -m oracle.make_golden_trainstep
"""
from __future__ import annotations

import os

import numpy as np
import torch

from buctd_b200 import synth
from buctd_b200.arch import HRNetArch
from buctd_b200.config import make_cfg
from oracle import condition_oracle as C
from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (variant, width, (W, H), cond, batch, att_modules)
CASES = {
    "train_step_plain_w32_128x96": ("plain", 32, (96, 128), "stacked", 4, (False, True, False, False)),
    "train_step_coam_w32_128x96": ("coam", 32, (96, 128), "stacked", 4, (False, True, False, False)),
}


def case_cfg(name):
    variant, width, size, cond, batch, att = CASES[name]
    return make_cfg(variant, width, size, cond=cond, att_modules=att), batch


def case_data(name, seed=4321):
    """(x [B, 3+Cc, H, W], target [B, K, H/4, W/4], target_weight [B, K, 1]) float32"""
    variant, width, (w, h), cond, batch, _ = CASES[name]
    rgb = synth.synth_rgb(batch, w, h, seed)
    x = rgb
    if variant != "plain":
        kp = synth.synth_keypoints(batch, 17, w, h, seed)
        c = np.stack([C.stacked_condition(kp[b], h, w) for b in range(batch)])
        x = np.concatenate([rgb, c.astype(np.float32)], 1)
    g = np.random.default_rng(seed + 1)
    K, hh, hw = 17, h // 4, w // 4
    tgt = np.zeros((batch, K, hh, hw), np.float32)
    yy, xx = np.mgrid[0:hh, 0:hw]
    for b in range(batch):
        for k in range(K):
            cx, cy = g.uniform(2, hw - 2), g.uniform(2, hh - 2)
            tgt[b, k] = np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 2.0 ** 2))
    tw = (g.uniform(size=(batch, K, 1)) < 0.8).astype(np.float32)
    return x.astype(np.float32), tgt, tw


def case_state_dict(name, seed=0):
    cfg, _ = case_cfg(name)
    return synth.synth_state_dict(HRNetArch(cfg), seed)


def main():
    if not R.available():
        raise SystemExit("reference tree not found at %s" % R.REF_ROOT)
    torch.set_num_threads(os.cpu_count() or 1)
    for name in CASES:
        cfg, batch = case_cfg(name)
        net = R.build_model(cfg, case_state_dict(name))
        net.train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.eval()
        x, tgt, tw = [torch.from_numpy(a) for a in case_data(name)]
        crit = R.joints_mse_loss(True)
        with R.cpu_cuda_shim():
            out = net(x)
        loss = crit(out, tgt, tw)
        net.zero_grad()
        loss.backward()
        res = {"loss": np.float64(loss.item()), "heatmaps": out.detach().numpy().astype(np.float32)}
        names, norms, heads = [], [], []
        for k, p in net.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            names.append(k)
            norms.append(float(g.double().norm()))
            h = np.zeros(12, np.float32)
            flat = g.reshape(-1)[:12].numpy()
            h[:flat.size] = flat
            heads.append(h)
        res["grad_names"] = np.array(names)
        res["grad_norms"] = np.array(norms, np.float64)
        res["grad_heads"] = np.stack(heads)
        sd = net.state_dict()
        for k in ("bn1", "layer1.0.bn1", "stage2.0.branches.1.0.bn2", "stage4.2.branches.0.3.bn2"):
            res["rm:" + k] = sd[k + ".running_mean"].numpy()
            res["rv:" + k] = sd[k + ".running_var"].numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **res)
        print(name, "loss %.6f" % loss.item(), "params", len(names), "max|hm| %.3f" % np.abs(res["heatmaps"]).max(),
              "|g| total %.4e" % float(np.sqrt((res["grad_norms"] ** 2).sum())))


if __name__ == "__main__":
    main()
