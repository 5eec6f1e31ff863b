"""TEST INFRASTRUCTURE -- the training step of the reference restated on top of oracle/hrnet_oracle.py (row a-T):

    model.train(); output = model(input); loss = JointsMSELoss(output, target, target_weight); loss.backward()
    (lib/core/function.py:113-138)

as a functional torch-fp32 forward in training mode (BatchNorm with batch statistics + running-stat update, attention
dropout with an explicit keep-mask) and torch.autograd for the gradients.  Pinned by tests/golden/train_step.npz, written
by oracle/make_golden_trainstep.py from the UNMODIFIED reference modules in train() mode (dropout switched off there: the
reference's mask comes from torch's RNG and cannot be shared; with the mask supplied the oracle's formula is the
definition of nn.Dropout).  Not a product path."""
from __future__ import annotations

import numpy as np
import torch

from oracle import hrnet_oracle as O


def drop_mask(seed: int, crop: int, rows: int, cols: int, p: float) -> np.ndarray:
    """keep[row][col] of one crop: the counter-based hash of include/buctd_b200.h ("Dropout"), uint32 arithmetic"""
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(cols, dtype=np.uint64)[None, :]
    M = np.uint64(0xFFFFFFFF)
    x = (np.uint64(seed) ^ ((r * np.uint64(0x9E3779B1) + c * np.uint64(0x85EBCA77) + np.uint64(crop) * np.uint64(0xC2B2AE3D)) & M)) & M
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    thr = np.uint64(min(int(p * 4294967296.0), 4294967295))
    return x >= thr


def joints_mse(output, target, target_weight, use_target_weight=True):
    """JointsMSELoss.forward, lib/core/loss.py:17-41 (closed form)"""
    B, K = output.shape[:2]
    o, t = output.reshape(B, K, -1), target.reshape(B, K, -1)
    if use_target_weight:
        w = target_weight.reshape(B, K, 1)
        o, t = o * w, t * w
    return 0.5 * ((o - t) ** 2).mean((0, 2)).sum() / K


def train_forward_backward(sd, cfg, x, target, target_weight, drop_p=0.0, seeds=None, use_target_weight=True):
    """sd: reference state_dict (float32).  seeds: {"<DAModule prefix>:chan" | ":pos" -> uint32 seed} for the hash dropout.
    Returns (loss, heat-maps, {param name -> gradient}, {bn prefix -> (running_mean, running_var) after the step})."""
    params = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point and "running_" not in k)
              for k, v in sd.items() if k.split(".")[-1] != "num_batches_tracked"}
    running = {}

    def drop(tag, att):
        seed = int(seeds[tag])
        B = att.shape[0]
        if tag.endswith(":chan"):
            C = att.shape[1]
            keep = np.stack([drop_mask(seed, b, C, C, drop_p) for b in range(B)])
        else:
            # position attention: the mask is indexed by PADDED pixel rows r = y (W + 1) + x of the act16 layout
            hw = att.shape[1]
            W = int(cfg["MODEL"]["IMAGE_SIZE"][0])
            H = int(cfg["MODEL"]["IMAGE_SIZE"][1])
            w = h = None
            for s in (4, 8, 16, 32):
                if (H // s) * (W // s) == hw:
                    h, w = H // s, W // s
            ridx = (np.arange(h)[:, None] * (w + 1) + np.arange(w)[None, :]).reshape(-1)
            rows = (h + 1) * (w + 1)
            keep = np.stack([drop_mask(seed, b, rows, rows, drop_p)[np.ix_(ridx, ridx)] for b in range(B)])
        return torch.from_numpy(keep.astype(np.float32))

    old = dict(O._CTX)
    O._CTX.update({"train": True, "running": running, "drop": drop if (drop_p > 0 and seeds) else None, "drop_p": float(drop_p)})
    try:
        out = O.forward(params, cfg, x)
    finally:
        O._CTX.clear()
        O._CTX.update(old)
    loss = joints_mse(out, target, target_weight, use_target_weight)
    names = [k for k, v in params.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    g = {k: (gr if gr is not None else torch.zeros_like(params[k])) for k, gr in zip(names, grads)}
    return float(loss), out.detach(), g, running
