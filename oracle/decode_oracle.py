"""TEST INFRASTRUCTURE -- numpy restatement of the reference heat-map decode and loss.
Not a product path: only tests/, smoke() and bench.py's CPU legs may import this.

* ``get_max_preds``    /root/reference/lib/core/inference.py:19-47
* ``get_final_preds``  /root/reference/lib/core/inference.py:51-87  (use_dark=False path)
* ``transform_preds`` / ``get_affine_transform(inv=1, rot=0)``  /root/reference/lib/utils/transforms.py:78-118
* ``joints_mse_loss``  /root/reference/lib/core/loss.py:17-41 (closed form)

Pinned by tests/test_oracle_vs_golden.py against outputs of the reference's own functions
(tests/golden/decode_*.npz, written by oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def get_max_preds(hm: np.ndarray):
    """hm [B,K,H,W] -> coords f32 [B,K,2] (x,y), maxvals f32 [B,K,1].
    argmax returns the FIRST maximal flat index; coords are zeroed where maxval <= 0."""
    assert isinstance(hm, np.ndarray), "batch_heatmaps should be numpy.ndarray"
    assert hm.ndim == 4, "batch_images should be 4-ndim"
    b, k, h, w = hm.shape
    flat = hm.reshape(b, k, -1)
    idx = np.argmax(flat, 2)
    maxvals = np.amax(flat, 2).reshape(b, k, 1)
    coords = np.empty((b, k, 2), np.float32)
    f = idx.astype(np.float32)
    coords[:, :, 0] = f % w
    coords[:, :, 1] = np.floor(f / w)
    coords *= (maxvals > 0.0).astype(np.float32)
    return coords, maxvals


def inverse_affine(center, scale, out_w: int, out_h: int):
    """The inv=1, rot=0 case of get_affine_transform: cv2.getAffineTransform of three point
    pairs reduces to a uniform similarity with factor f = 200*scale[0]/out_w (only scale[0] is
    used, transforms.py:94-101):  x_img = cx + (x - out_w/2) f ;  y_img = cy + (y - out_h/2) f.
    Returns the 2x3 float64 matrix the reference would obtain."""
    c = np.asarray(center, np.float32).astype(np.float64)
    src_w = np.float32(np.asarray(scale, np.float32)[0] * np.float32(200.0)).astype(np.float64)
    f = src_w / float(out_w)
    return np.array([[f, 0.0, c[0] - f * out_w * 0.5],
                     [0.0, f, c[1] - f * out_h * 0.5]], np.float64)


def get_final_preds(post_process: bool, hm: np.ndarray, center, scale):
    coords, maxvals = get_max_preds(hm)
    b, k, h, w = hm.shape
    if post_process:
        px = np.floor(coords[:, :, 0] + 0.5).astype(np.int64)
        py = np.floor(coords[:, :, 1] + 0.5).astype(np.int64)
        ok = (1 < px) & (px < w - 1) & (1 < py) & (py < h - 1)
        bi, ki = np.nonzero(ok)
        x, y = px[bi, ki], py[bi, ki]
        dx = hm[bi, ki, y, x + 1] - hm[bi, ki, y, x - 1]
        dy = hm[bi, ki, y + 1, x] - hm[bi, ki, y - 1, x]
        # coords is float32; np.sign(float32 diff) * .25 is added in float32 (exact: quarter steps)
        coords[bi, ki, 0] += (np.sign(dx) * 0.25).astype(np.float32)
        coords[bi, ki, 1] += (np.sign(dy) * 0.25).astype(np.float32)
    preds = np.empty_like(coords)
    for i in range(b):
        t = inverse_affine(center[i], scale[i], w, h)
        pts = np.concatenate([coords[i].astype(np.float64), np.ones((k, 1))], 1)
        preds[i] = (pts @ t.T).astype(np.float32)
    return preds, maxvals


def joints_mse_loss(output: np.ndarray, target: np.ndarray, target_weight: np.ndarray,
                    use_target_weight: bool = True) -> float:
    """0.5/(K*B*HW) * sum (w*(pred-gt))^2 ; returns a python float (float64 accumulation)."""
    b, k = output.shape[:2]
    p = output.reshape(b, k, -1).astype(np.float64)
    g = target.reshape(b, k, -1).astype(np.float64)
    if use_target_weight:
        w = target_weight.reshape(b, k, 1).astype(np.float64)
        p, g = p * w, g * w
    return float(0.5 * np.mean((p - g) ** 2, axis=(0, 2)).sum() / k)
