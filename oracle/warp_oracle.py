"""TEST INFRASTRUCTURE -- numpy restatement of the crop pipeline the reference runs in its DataLoader workers
(SURVEY.md section 8, row f-2).  Not a product path: only tests/ may import this.

* ``get_affine_transform``  /root/reference/lib/utils/transforms.py:86-118 (+ get_dir, get_3rd_point :134-150); the 2x3
  matrix itself is ``cv2.getAffineTransform`` -- third party, opencv-python (reference pin 4.5.5.64; 4.13.0 here): the
  exact affine map through three point pairs, solved in float64.
* ``invert_affine``         cv::invertAffineTransform as warpAffine applies it to the forward matrix (OpenCV
  modules/imgproc/src/imgwarp.cpp, warpAffine): float64, D = 1/det, no pivoting.
* ``warp_affine_u8``        cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR) for 8-bit images, BORDER_CONSTANT 0:
  OpenCV's fixed-point path (imgwarp.cpp WarpAffineInvoker + remapBilinear<FixedPtCast<int, uchar, 15>>): coordinates
  in 1/1024 px (round half even) + 16, >> 5 -> 1/32 px; tap weights (32-fx)(32-fy)*32 ...; (sum + 2^14) >> 15.
* ``to_tensor_normalize``   torchvision ToTensor + Normalize as tools/train.py:171-177 composes them.

Pinned against cv2.warpAffine / cv2.getAffineTransform themselves (bit for bit) and the reference's
get_affine_transform in tests/test_warp_oracle.py, and against tests/golden/crop_warp.npz.
"""
from __future__ import annotations

import numpy as np


def get_dir(src_point, rot_rad):
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [src_point[0] * cs - src_point[1] * sn, src_point[0] * sn + src_point[1] * cs]


def get_3rd_point(a, b):
    direct = a - b
    return b + np.array([-direct[1], direct[0]], dtype=np.float32)


def affine_from_points(src, dst):
    """the affine map with dst_i = M [src_i; 1] for three point pairs, float64 (what cv2.getAffineTransform solves)"""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    A = np.concatenate([src, np.ones((3, 1))], 1)
    return np.linalg.solve(A, dst).T


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0, solver=None):
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    src_dir = get_dir([0, src_w * -0.5], rot_rad)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2:, :] = get_3rd_point(src[0, :], src[1, :])
    dst[2:, :] = get_3rd_point(dst[0, :], dst[1, :])
    solve = solver if solver is not None else affine_from_points
    return solve(np.float32(dst), np.float32(src)) if inv else solve(np.float32(src), np.float32(dst))


def invert_affine(M):
    m = np.array(M, np.float64).reshape(6).copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def warp_affine_u8(img, M, dsize):
    W, H = dsize
    m = invert_affine(M).reshape(6)
    x, y = np.arange(W), np.arange(H)
    adelta = np.rint(m[0] * x * 1024).astype(np.int64)
    bdelta = np.rint(m[3] * x * 1024).astype(np.int64)
    X0 = np.rint((m[1] * y + m[2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((m[4] * y + m[5]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy, fx, fy = X >> 5, Y >> 5, X & 31, Y & 31
    h, w = img.shape[:2]

    def px(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64) * ok[..., None]

    w00, w01 = ((32 - fx) * (32 - fy) * 32)[..., None], (fx * (32 - fy) * 32)[..., None]
    w10, w11 = ((32 - fx) * fy * 32)[..., None], (fx * fy * 32)[..., None]
    out = (px(sy, sx) * w00 + px(sy, sx + 1) * w01 + px(sy + 1, sx) * w10 + px(sy + 1, sx + 1) * w11 + (1 << 14)) >> 15
    return out.astype(np.uint8)


def to_tensor_normalize(img_u8, mean, std):
    """HWC uint8 -> CHW float32: x / 255, then (x - mean) / std, every step in float32"""
    x = img_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return (x - np.asarray(mean, np.float32)[:, None, None]) / np.asarray(std, np.float32)[:, None, None]
