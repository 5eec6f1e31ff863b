"""TEST INFRASTRUCTURE -- numpy restatement of the reference code either side of the forward in the
validation / training loops.  Not a product path: only tests/, smoke() and bench.py's CPU legs may import this.

* ``flip_back``            /root/reference/lib/utils/transforms.py:16-30
* ``flip_hm_stacked``      /root/reference/lib/utils/transforms.py:49-55 (stacked-condition branch of flip_hm)
* ``flip_test_input``      /root/reference/lib/core/function.py:213-217
* ``flip_test_merge``      /root/reference/lib/core/function.py:226-236 (flip_back, SHIFT_HEATMAP, average)
* ``joints_mse_grad``      autograd gradient of JointsMSELoss, /root/reference/lib/core/loss.py:17-41 (closed form)
* ``adam_step``            torch.optim.Adam as built by get_optimizer, /root/reference/lib/utils/utils.py:268-272
                           (third-party: torch; single-tensor algorithm of torch/optim/adam.py, torch 2.11)
* ``generate_target``      /root/reference/lib/dataset/JointsDataset.py:397-453 (Gaussian training targets + weights)
* ``accuracy``             /root/reference/lib/core/evaluate.py:14-70 (PCK on heat-map arg-maxes; calc_dists, dist_acc)
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


def gaussian_stamp(sigma) -> np.ndarray:
    """The (6 sigma + 1)^2 un-normalised Gaussian of generate_target, float32 like the reference's arange
    (lib/dataset/JointsDataset.py:428-433)."""
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))


def generate_target(joints, joints_vis, image_size, heatmap_size, sigma, joints_weight=None):
    """joints, joints_vis [K,3] (one crop) -> (target f32 [K,H,W], target_weight f32 [K,1]);
    lib/dataset/JointsDataset.py:397-453 with target_type 'gaussian'."""
    K = joints.shape[0]
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    target_weight = np.ones((K, 1), dtype=np.float32)
    target_weight[:, 0] = joints_vis[:, 0]
    target = np.zeros((K, H, W), dtype=np.float32)
    tmp_size = sigma * 3
    g = gaussian_stamp(sigma)
    feat_stride = np.asarray(image_size) / np.asarray(heatmap_size)
    for k in range(K):
        mu_x = int(joints[k][0] / feat_stride[0] + 0.5)
        mu_y = int(joints[k][1] / feat_stride[1] + 0.5)
        ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
        br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            target_weight[k] = 0
            continue
        g_x = max(0, -ul[0]), min(br[0], W) - ul[0]
        g_y = max(0, -ul[1]), min(br[1], H) - ul[1]
        img_x = max(0, ul[0]), min(br[0], W)
        img_y = max(0, ul[1]), min(br[1], H)
        if target_weight[k] > 0.5:
            target[k][img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    if joints_weight is not None:
        target_weight = np.multiply(target_weight, joints_weight)
    return target, target_weight


def accuracy(output, target, thr=0.5):
    """(acc float64 [K+1], avg_acc, cnt, pred) of lib/core/evaluate.py:40-70 (hm_type 'gaussian')."""
    from oracle import decode_oracle as D
    pred, _ = D.get_max_preds(output)
    tgt, _ = D.get_max_preds(target)
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    B, K = pred.shape[0], pred.shape[1]
    dists = np.zeros((K, B))
    p32, t32 = pred.astype(np.float32), tgt.astype(np.float32)
    for n in range(B):
        for c in range(K):
            if t32[n, c, 0] > 1 and t32[n, c, 1] > 1:
                dists[c, n] = np.linalg.norm(p32[n, c, :] / norm[n] - t32[n, c, :] / norm[n])
            else:
                dists[c, n] = -1
    acc = np.zeros(K + 1)
    avg, cnt = 0.0, 0
    for i in range(K):
        cal = np.not_equal(dists[i], -1)
        acc[i + 1] = np.less(dists[i][cal], thr).sum() * 1.0 / cal.sum() if cal.sum() > 0 else -1
        if acc[i + 1] >= 0:
            avg += acc[i + 1]
            cnt += 1
    avg = avg / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg
    return acc, avg, cnt, pred
