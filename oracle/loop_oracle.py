"""TEST INFRASTRUCTURE -- numpy restatement of the reference code either side of the forward in the
validation / training loops.  Not a product path: only tests/, smoke() and bench.py's CPU legs may import this.

* ``flip_back``            /root/reference/lib/utils/transforms.py:16-30
* ``flip_hm_stacked``      /root/reference/lib/utils/transforms.py:49-55 (stacked-condition branch of flip_hm)
* ``flip_test_input``      /root/reference/lib/core/function.py:213-217
* ``flip_test_merge``      /root/reference/lib/core/function.py:226-236 (flip_back, SHIFT_HEATMAP, average)
* ``joints_mse_grad``      autograd gradient of JointsMSELoss, /root/reference/lib/core/loss.py:17-41 (closed form)
* ``adam_step``            torch.optim.Adam as built by get_optimizer, /root/reference/lib/utils/utils.py:268-272
                           (third-party: torch; single-tensor algorithm of torch/optim/adam.py, torch 2.11)
* ``refine_condition``     the synthetic fixed-crop refinement loop of BASELINE.json configs[3]

Pinned by tests/test_oracle_vs_golden.py against the reference's own flip_back / flip_hm / JointsMSELoss autograd and
torch.optim.Adam (tests/golden/loop_ops.npz, written by oracle/make_golden_loop.py).
"""
from __future__ import annotations

import numpy as np


def _swap_pairs(a: np.ndarray, pairs) -> np.ndarray:
    a = a.copy()
    for p0, p1 in pairs:                       # sequential swaps, exactly like the reference loop
        tmp = a[:, p0].copy()
        a[:, p0] = a[:, p1]
        a[:, p1] = tmp
    return a


def flip_back(output_flipped: np.ndarray, matched_parts) -> np.ndarray:
    assert output_flipped.ndim == 4, "output_flipped should be [batch_size, num_joints, height, width]"
    return _swap_pairs(output_flipped[:, :, :, ::-1], matched_parts)


def flip_hm_stacked(heatmap: np.ndarray, matched_parts) -> np.ndarray:
    return _swap_pairs(heatmap[:, :, :, ::-1], matched_parts)


def pairs_to_perm(n: int, pairs) -> np.ndarray:
    """perm with out[:, c] = in[:, perm[c]] equal to the sequential pair swaps"""
    idx = np.arange(n)[None, :, None, None].astype(np.int64)
    return _swap_pairs(idx, pairs)[0, :, 0, 0]


def flip_test_input(x: np.ndarray, flip_pairs, stacked: bool) -> np.ndarray:
    rgb = x[:, :3, :, ::-1]
    cond = x[:, 3:]
    if cond.shape[1] == 0:
        return np.ascontiguousarray(rgb)
    cond = flip_hm_stacked(cond, flip_pairs) if stacked else cond[:, :, :, ::-1]
    return np.ascontiguousarray(np.concatenate([rgb, cond], 1))


def flip_test_merge(output: np.ndarray, output_flipped: np.ndarray, flip_pairs, shift_heatmap: bool) -> np.ndarray:
    f = flip_back(output_flipped, flip_pairs).copy()
    if shift_heatmap:
        f[:, :, :, 1:] = f.copy()[:, :, :, 0:-1]
    return ((output + f) * np.float32(0.5)).astype(np.float32)


def joints_mse_grad(output, target, target_weight, use_target_weight=True) -> np.ndarray:
    b, k = output.shape[:2]
    hw = int(np.prod(output.shape[2:]))
    w = target_weight.reshape(b, k, *([1] * (output.ndim - 2))).astype(np.float64) if use_target_weight else 1.0
    return (w * w * (output.astype(np.float64) - target.astype(np.float64)) / (k * b * hw)).astype(np.float32)


def adam_step(p, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """float32 arrays in, (p, m, v) float32 out; scalar arithmetic in Python doubles like torch"""
    f = np.float32
    m = m + (g - m) * f(1.0 - beta1)
    v = v * f(beta2) + f(1.0 - beta2) * g * g
    bc1, bc2 = 1.0 - beta1 ** step, 1.0 - beta2 ** step
    denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


def refine_condition(coords_hm: np.ndarray, stride: float = 4.0) -> np.ndarray:
    """heat-map coordinates (x, y) of pass n -> crop-pixel keypoints rasterised for pass n+1 (fixed crop)"""
    return (coords_hm * np.float32(stride)).astype(np.float32)
