"""Crops on the device: the reference warps every person crop with cv2.warpAffine in its
DataLoader workers, converts it with ToTensor + Normalize and copies it host -> device
(/root/reference/lib/dataset/JointsDataset.py:262-288, tools/train.py:171-177, tools/inference.py:86-107).  Here the
8-bit source images live in device memory and one kernel writes the normalised crops straight into the model's input
tensor (bit-identical to cv2 + torchvision); bbox -> (center, scale) is ``buctd_b200.results.xywh2cs``."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # tools/train.py:171-173
IMAGENET_STD = (0.229, 0.224, 0.225)


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """lib/utils/transforms.py:86-118, same signature and result (the matrix comes from cv2.getAffineTransform, exactly
    as in the reference: 18 flops per crop on the host)."""
    import cv2
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):
        d = a - b
        return b + np.array([-d[1], d[0]], dtype=np.float32)
    src[2:, :] = third(src[0, :], src[1, :])
    dst[2:, :] = third(dst[0, :], dst[1, :])
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def affine_transform(pt, t):
    """lib/utils/transforms.py:121-124"""
    new_pt = np.array([pt[0], pt[1], 1.]).T
    return np.dot(t, new_pt)[:2]


def _invert(trans):
    # cv::invertAffineTransform as cv2.warpAffine applies it to the forward matrix (float64, D = 1 / det)
    m = np.array(trans, np.float64).reshape(6).copy()
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
    return m


def warp_crops(images, trans, output_size, out=None, c_off=0, out_channels=3, mean=IMAGENET_MEAN, std=IMAGENET_STD,
               bgr_to_rgb=False, want_u8=False):
    """images: list of B CUDA uint8 tensors [h_i, w_i, 3] (one per crop; the same tensor may appear many times);
    trans: [B, 2, 3] forward matrices (``get_affine_transform(c, s, r, image_size)``); output_size = (w, h) like
    cfg.MODEL.IMAGE_SIZE.  Writes planes [c_off, c_off + 3) of ``out`` (float32 [B, out_channels, h, w], allocated if
    None) and returns it (plus the 8-bit crops [B, h, w, 3] with want_u8)."""
    B = len(images)
    W, H = int(output_size[0]), int(output_size[1])
    if B and not images[0].is_cuda:
        raise L.BuctdError("warp_crops needs CUDA images; there is no CPU path")
    dev = images[0].device if B else torch.device("cuda", torch.cuda.current_device())
    for im in images:
        assert im.dtype == torch.uint8 and im.dim() == 3 and im.shape[2] == 3 and im.stride(2) == 1 and im.stride(1) == 3
    if out is None:
        out = torch.zeros(B, out_channels, H, W, dtype=torch.float32, device=dev)
    assert out.is_contiguous() and out.shape[0] == B and out.shape[2] == H and out.shape[3] == W
    if B == 0:
        return (out, torch.empty(0, H, W, 3, dtype=torch.uint8, device=dev)) if want_u8 else out
    inv = np.stack([_invert(t) for t in np.asarray(trans, np.float64).reshape(B, 6)])
    ptrs = torch.tensor([im.data_ptr() for im in images], dtype=torch.int64).to(dev)
    hs = torch.tensor([im.shape[0] for im in images], dtype=torch.int32).to(dev)
    ws = torch.tensor([im.shape[1] for im in images], dtype=torch.int32).to(dev)
    strides = torch.tensor([im.stride(0) for im in images], dtype=torch.int64).to(dev)
    inv_t = torch.from_numpy(np.ascontiguousarray(inv)).to(dev)
    u8 = torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev) if want_u8 else None
    a = L.CropWarpArgs()
    a.batch, a.h, a.w, a.out_channels, a.c_off, a.swap_rb = B, H, W, out.shape[1], int(c_off), int(bool(bgr_to_rgb))
    for i in range(3):
        a.mean[i], a.std[i] = float(np.float32(mean[i])), float(np.float32(std[i]))
    a.images, a.img_h, a.img_w, a.img_stride = ptrs.data_ptr(), hs.data_ptr(), ws.data_ptr(), strides.data_ptr()
    a.inv_trans, a.out, a.out_u8 = inv_t.data_ptr(), out.data_ptr(), (u8.data_ptr() if u8 is not None else None)
    with torch.cuda.device(dev):
        L.Context.get(dev.index).call("buctd_crop_warp", a, torch.cuda.current_stream(dev).cuda_stream)
    warp_crops._keep = (images, ptrs, hs, ws, strides, inv_t)
    return (out, u8) if want_u8 else out
