"""TEST INFRASTRUCTURE -- writes tests/golden/loop_ops.npz from the UNMODIFIED reference functions
(lib/utils/transforms.flip_back / flip_hm, lib/core/loss.JointsMSELoss + torch autograd) and torch.optim.Adam
(the optimizer the reference's get_optimizer constructs) on seeded inputs.  Run in the build container:

    python -m oracle.make_golden_loop
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]   # lib/dataset/coco.py:66-67


class _FakeDataset:
    flip_pairs = COCO_FLIP_PAIRS


def main():
    if not R.available():
        sys.exit("reference tree not found at %s" % R.REF_ROOT)
    T = R._import()["transforms"]
    g = np.random.default_rng(23)
    out = {"flip_pairs": np.array(COCO_FLIP_PAIRS, np.int32)}
    # flip_back on heat-maps, flip_hm on a stacked condition, and the function.py merge with / without SHIFT_HEATMAP
    hm = g.standard_normal((3, 17, 12, 9)).astype(np.float32)
    hm_f = g.standard_normal((3, 17, 12, 9)).astype(np.float32)
    out["hm"], out["hm_flipped"] = hm, hm_f
    out["flip_back"] = np.ascontiguousarray(T.flip_back(hm_f.copy(), COCO_FLIP_PAIRS))
    x = g.standard_normal((2, 20, 16, 12)).astype(np.float32)
    out["x"] = x.copy()
    xt = torch.from_numpy(x)
    # .clone(): in the reference the input is a CUDA tensor, so flip_hm's .cpu().numpy() is a copy; on a CPU tensor it
    # would alias (and swap the channels of) the input itself
    flipped = torch.cat((xt[:, :3].flip(3), T.flip_hm(xt[:, 3:].clone(), _FakeDataset(), None, None)), axis=1)   # function.py:215
    out["x_flipped"] = flipped.numpy().astype(np.float32)
    for shift in (0, 1):
        of = torch.from_numpy(np.ascontiguousarray(T.flip_back(hm_f.copy(), COCO_FLIP_PAIRS)).copy())
        if shift:
            of[:, :, :, 1:] = of.clone()[:, :, :, 0:-1]                                                  # function.py:232-234
        out["merge_shift%d" % shift] = ((torch.from_numpy(hm) + of) * 0.5).numpy()
    # JointsMSELoss gradient through autograd
    o = g.standard_normal((4, 17, 16, 12)).astype(np.float32)
    t = g.uniform(0, 1, (4, 17, 16, 12)).astype(np.float32)
    wgt = (g.uniform(size=(4, 17, 1)) < 0.8).astype(np.float32)
    out["loss_output"], out["loss_target"], out["loss_weight"] = o, t, wgt
    for utw in (0, 1):
        ot = torch.from_numpy(o).requires_grad_(True)
        loss = R.joints_mse_loss(bool(utw))(ot, torch.from_numpy(t), torch.from_numpy(wgt))
        loss.backward()
        out["loss_utw%d" % utw] = np.float32(loss.item())
        out["grad_utw%d" % utw] = ot.grad.numpy().astype(np.float32)
    # three Adam steps (torch.optim.Adam(params, lr=1e-3): lib/utils/utils.py:268-272)
    p0 = g.standard_normal(1003).astype(np.float32)
    grads = [(g.standard_normal(1003) * s).astype(np.float32) for s in (1.0, 0.1, 3.0)]
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=1e-3)
    out["adam_p0"] = p0
    for i, gr in enumerate(grads):
        p.grad = torch.from_numpy(gr.copy())
        opt.step()
        out["adam_g%d" % i] = gr
        out["adam_p%d" % (i + 1)] = p.detach().numpy().copy()
    st = opt.state[p]
    out["adam_m3"], out["adam_v3"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN, "loop_ops.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "loop_ops.npz"))


if __name__ == "__main__":
    main()
