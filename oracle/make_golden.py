"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, fp32) on seeded synthetic inputs.  Run in the build container:

    python -m oracle.make_golden

The fixtures hold only outputs (and tiny hand-made decode inputs); the inputs and weights are
regenerated from seeds by buctd_b200.synth (numpy PCG64, machine independent) and
oracle.condition_oracle, see ``case_inputs``.  Changing buctd_b200.synth.SCALES or the seeds
invalidates the fixtures: regenerate and commit them together.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from buctd_b200 import synth
from buctd_b200.arch import HRNetArch
from buctd_b200.config import make_cfg
from oracle import condition_oracle as C
from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (variant, width, (W,H), cond, batch)
CASES = {
    "plain_w32_256x192": ("plain", 32, (192, 256), "stacked", 1),        # BASELINE.json configs[0]
    "prenet_w48_384x288": ("prenet", 48, (288, 384), "colored", 1),      # configs[1] (B=1 sample)
    "coam_w48_384x288_stacked": ("coam", 48, (288, 384), "stacked", 2),  # configs[2] (B=2 sample)
    "coam_w48_384x288_colored": ("coam", 48, (288, 384), "colored", 1),
    "coam_w32_128x96_stacked": ("coam", 32, (96, 128), "stacked", 2),    # small, for fast tests
    "coam_w32_128x96_plaincond": ("coam", 32, (96, 128), "plain", 1),    # 1-channel condition path
    # every CoAM placement (MODEL.ATT_MODULES; pose_hrnet_coam.py:521,535,554,561): stage1_att (2 branches), the
    # reference's default [F,F,T,T] (stage3_att on 4 branches incl. the 1/32 scale + stage4_att on branch 0), stage4_att alone
    "coam_w32_128x96_att1000": ("coam", 32, (96, 128), "stacked", 2, (True, False, False, False)),
    "coam_w32_128x96_att0011": ("coam", 32, (96, 128), "stacked", 2, (False, False, True, True)),
    "coam_w32_128x96_att0001": ("coam", 32, (96, 128), "colored", 1, (False, False, False, True)),
    "coam_w48_384x288_att0011": ("coam", 48, (288, 384), "stacked", 1, (False, False, True, True)),
}


def case_cfg(name):
    variant, width, size, cond, batch = CASES[name][:5]
    att = CASES[name][5] if len(CASES[name]) > 5 else (False, True, False, False)
    return make_cfg(variant, width, size, cond=cond, att_modules=att), batch


def case_inputs(name, seed=1234):
    """x [B,3+Cc,H,W] float32 exactly as the golden run saw it."""
    variant, width, (w, h), cond, batch = CASES[name][:5]
    rgb = synth.synth_rgb(batch, w, h, seed)
    if variant == "plain":
        return rgb
    kp = synth.synth_keypoints(batch, 17, w, h, seed)
    if cond == "stacked":
        c = np.stack([C.stacked_condition(kp[b], h, w) for b in range(batch)])
    elif cond == "colored":
        c = np.stack([C.colored_condition(kp[b], h, w, C.default_colors()) for b in range(batch)])
    else:
        c = np.stack([C.plain_condition(kp[b], h, w) for b in range(batch)])
    return np.concatenate([rgb, c.astype(np.float32)], 1)


def case_state_dict(name, seed=0):
    cfg, _ = case_cfg(name)
    return synth.synth_state_dict(HRNetArch(cfg), seed)


def decode_cases():
    """Hand-made heat-maps for the decode edge cases the reference's arithmetic has."""
    g = np.random.default_rng(7)
    b, k, h, w = 3, 17, 24, 18
    hm = g.standard_normal((b, k, h, w)).astype(np.float32)
    hm[0, 0] = -1.0                       # all negative & all equal: idx 0, coords masked to 0
    hm[0, 1] = 0.0                        # all zero: maxval 0 -> masked
    hm[0, 2] = 0.0; hm[0, 2, 5, 7] = 0.5; hm[0, 2, 9, 3] = 0.5   # tie: first flat index wins
    hm[0, 3] = 0.0; hm[0, 3, 0, 0] = 1.0                        # corner: no quarter offset
    hm[0, 4] = 0.0; hm[0, 4, 1, 1] = 1.0                        # px=1: strict 1<px fails
    hm[0, 5] = 0.0; hm[0, 5, 2, 2] = 1.0; hm[0, 5, 2, 3] = 0.3; hm[0, 5, 3, 2] = 0.3  # +,+
    hm[0, 6] = 0.0; hm[0, 6, h - 2, w - 2] = 1.0                # px=W-2: px<W-1 fails? (W-2<W-1 ok) boundary
    hm[0, 7] = 0.0; hm[0, 7, h - 3, w - 3] = 1.0; hm[0, 7, h - 3, w - 4] = 0.2        # -x
    hm[0, 8] = 0.0; hm[0, 8, 10, 10] = 1.0                      # equal neighbours: sign 0
    hm[0, 9] = -5.0; hm[0, 9, 4, 4] = -0.1                      # negative max: coords masked, maxval kept
    hm[0, 10] = 0.0; hm[0, 10, h - 1, w - 1] = 2.0              # last element
    center = np.array([[9.3, 12.1], [100.5, 80.25], [144.0, 192.0]], np.float32)
    scale = np.array([[0.5, 0.7], [1.44, 1.92], [1.7, 2.3]], np.float32)
    return hm, center, scale


def main():
    if not R.available():
        sys.exit("reference tree not found at %s" % R.REF_ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name in CASES:
        if only and name not in only:
            continue
        cfg, batch = case_cfg(name)
        sd = case_state_dict(name)
        net = R.build_model(cfg, sd)
        x = torch.from_numpy(case_inputs(name))
        hm = R.forward(net, x).numpy().astype(np.float32)
        w, h = cfg.MODEL.IMAGE_SIZE
        center, scale = synth.synth_center_scale(batch, w, h)
        cfg.TEST.POST_PROCESS = True
        preds, maxvals = R.get_final_preds(cfg, hm.copy(), center, scale)
        coords, _ = R.get_max_preds(hm.copy())
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), heatmaps=hm, preds=preds.astype(np.float32),
                            maxvals=maxvals.astype(np.float32), coords=coords.astype(np.float32),
                            center=center, scale=scale)
        print(name, hm.shape, "max|hm| %.4f" % np.abs(hm).max())
    if only:
        return

    hm, center, scale = decode_cases()
    cfg = make_cfg("plain", 32, (72, 96))
    out = {"heatmaps": hm, "center": center, "scale": scale}
    for pp in (False, True):
        cfg.TEST.POST_PROCESS = pp
        preds, maxvals = R.get_final_preds(cfg, hm.copy(), center, scale)
        out["preds_pp%d" % pp] = preds.astype(np.float32)
        out["maxvals"] = maxvals.astype(np.float32)
    coords, _ = R.get_max_preds(hm.copy())
    out["coords"] = coords.astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN, "decode_edge_cases.npz"), **out)

    # JointsMSELoss (lib/core/loss.py:17-41) on seeded inputs, both weighting modes
    g = np.random.default_rng(11)
    o = g.standard_normal((4, 17, 16, 12)).astype(np.float32)
    t = g.uniform(0, 1, (4, 17, 16, 12)).astype(np.float32)
    wgt = (g.uniform(size=(4, 17, 1)) < 0.8).astype(np.float32)
    res = {"output": o, "target": t, "target_weight": wgt}
    for utw in (False, True):
        crit = R.joints_mse_loss(utw)
        res["loss_utw%d" % utw] = np.float32(crit(torch.from_numpy(o), torch.from_numpy(t), torch.from_numpy(wgt)).item())
    np.savez_compressed(os.path.join(GOLDEN, "joints_mse.npz"), **res)
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
