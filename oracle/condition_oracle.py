"""TEST INFRASTRUCTURE -- numpy restatement of the reference's condition rasteriser.

The reference rasterises bottom-up keypoint proposals inside ``JointsDataset`` (un-importable
here: it needs pycocotools/json_tricks/matplotlib), so this file restates it:

* ``generate_heatmap``            /root/reference/lib/dataset/JointsDataset.py:457-463
* ``get_stacked_condition``       :471-498   one channel per keypoint, each normalised on its own
* ``get_condition_image``         :500-516   one channel, all keypoints, repeated x3, ``.astype(int)``
* ``get_condition_image_colored`` :519-543   3 channels, later keypoints overwrite, normalised by
                                             the global max (standalone copy tools/inference.py:163-185)

``cv2.GaussianBlur(ksize=(15,15), sigmaX=0)`` means sigma = 0.3*((15-1)*0.5 - 1) + 0.8 = 2.6 with
BORDER_REFLECT_101; ``gaussian_blur_15`` writes that out as a separable 15-tap filter in float64
and tests/test_condition_oracle.py checks it against cv2 itself (max err ~1e-13).
Not a product path: only tests/, smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np

KSIZE = 15
SIGMA = 0.3 * ((KSIZE - 1) * 0.5 - 1) + 0.8          # 2.6  (OpenCV's rule for sigma<=0)


def gaussian_kernel_15() -> np.ndarray:
    r = np.arange(KSIZE, dtype=np.float64) - (KSIZE - 1) / 2
    k = np.exp(-(r * r) / (2.0 * SIGMA * SIGMA))
    return k / k.sum()


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def gaussian_blur_15(img: np.ndarray) -> np.ndarray:
    """img [H,W] or [H,W,C] float64 -> same shape; separable 15x15, sigma 2.6, reflect-101."""
    k = gaussian_kernel_15()
    h, w = img.shape[:2]
    r = np.arange(-7, 8)
    xi = _reflect101(np.arange(w)[:, None] + r[None, :], w)       # [W,15]
    yi = _reflect101(np.arange(h)[:, None] + r[None, :], h)       # [H,15]
    out = np.tensordot(img[:, xi], k, axes=([2], [0])) if img.ndim == 2 else \
        np.einsum("hwtc,t->hwc", img[:, xi], k)
    out = np.tensordot(out[yi], k, axes=([1], [0])) if img.ndim == 2 else \
        np.einsum("htwc,t->hwc", out[yi], k)
    return out


def _generate_heatmap(hm: np.ndarray) -> np.ndarray:
    hm = gaussian_blur_15(hm)
    am = np.amax(hm)
    if am == 0:
        return hm
    return hm / (am / 255)


def stacked_condition(kpts: np.ndarray, h: int, w: int) -> np.ndarray:
    """kpts [K,>=2] -> [K,H,W] float64, values 0..255 (channel-first; the reference returns
    [H,W,K] and the dataset later transposes it)."""
    kp = np.array(kpts).astype(int)
    out = np.zeros((len(kp), h, w), np.float64)
    for i, k in enumerate(kp):
        if 0 < k[0] < w and 0 < k[1] < h:
            m = np.zeros((h, w), np.float64)
            m[k[1] - 1][k[0] - 1] = 255
            out[i] = _generate_heatmap(m)
    return out


def plain_condition(kpts: np.ndarray, h: int, w: int) -> np.ndarray:
    """[3,H,W] int64: one blurred impulse map, repeated three times, truncated to int."""
    kp = np.array(kpts).astype(int)
    m = np.zeros((h, w), np.float64)
    for k in kp:
        if 0 < k[0] < w and 0 < k[1] < h:
            m[k[1] - 1][k[0] - 1] = 255
    hm = _generate_heatmap(m)
    return np.repeat(hm[None], 3, axis=0).astype(int)


def colored_condition(kpts: np.ndarray, h: int, w: int, colors) -> np.ndarray:
    """[3,H,W] float64 (channel-first view of the reference's [H,W,3])."""
    kp = np.array(kpts).astype(int)
    m = np.zeros((h, w, 3), np.float64)
    for color, k in zip(colors, kp):
        if 0 < k[0] < w and 0 < k[1] < h:
            m[k[1] - 1][k[0] - 1] = color
    return np.moveaxis(_generate_heatmap(m), -1, 0)


# COCO keypoint colours: the data table at /root/reference/lib/dataset/coco.py:54
COCO_KPT_COLORS = [[245, 59, 59], [249, 104, 25], [253, 183, 15], [233, 245, 41], [162, 252, 32],
                   [84, 247, 34], [31, 252, 57], [20, 246, 126], [5, 249, 206], [52, 215, 249],
                   [33, 136, 252], [11, 39, 248], [93, 46, 249], [156, 29, 244], [235, 49, 247],
                   [245, 47, 187], [253, 44, 117]]


def default_colors(n: int = 17):
    return [COCO_KPT_COLORS[i % len(COCO_KPT_COLORS)] for i in range(n)]
