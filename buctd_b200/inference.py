"""Drop-in for the reference's ``core.inference``: ``get_max_preds`` / ``get_final_preds`` with the
reference signatures (NumPy in, NumPy out; /root/reference/lib/core/inference.py:19-87), computed by
one fused CUDA kernel (argmax + quarter-pixel refine + inverse crop transform), plus a
device-tensor fast path (``decode_device``) that avoids the D2H copy of the heat-maps the
reference pays in its validation loop (lib/core/function.py:260-261)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return t.data_ptr() if t is not None else None


def decode_device(heatmaps: torch.Tensor, center=None, scale=None, post_process: bool = True,
                  want_coords: bool = False):
    """heatmaps: CUDA float32 [B,K,H,W] (contiguous).  center/scale: CUDA float32 [B,2] or None
    (None: predictions stay in heat-map coordinates).  Returns (preds [B,K,2], maxvals [B,K,1]
    [, coords [B,K,2], idx int32 [B,K]]) as CUDA tensors; nothing is synchronised."""
    if not heatmaps.is_cuda:
        raise L.BuctdError("decode_device needs a CUDA tensor; there is no CPU path")
    assert heatmaps.dim() == 4, "batch_images should be 4-ndim"
    hm = heatmaps.contiguous().float()
    B, K, H, W = hm.shape
    dev = hm.device
    preds = torch.empty(B, K, 2, dtype=torch.float32, device=dev)
    maxvals = torch.empty(B, K, 1, dtype=torch.float32, device=dev)
    coords = torch.empty(B, K, 2, dtype=torch.float32, device=dev) if want_coords else None
    idx = torch.empty(B, K, dtype=torch.int32, device=dev) if want_coords else None
    if center is not None:
        center = torch.as_tensor(center, dtype=torch.float32).to(dev).contiguous()
        scale = torch.as_tensor(scale, dtype=torch.float32).to(dev).contiguous()
    a = L.DecodeArgs()
    a.batch, a.k, a.h, a.w, a.post_process = B, K, H, W, int(bool(post_process))
    a.heatmaps, a.center, a.scale = hm.data_ptr(), _ptr(center), _ptr(scale)
    a.preds, a.maxvals, a.coords, a.idx = preds.data_ptr(), maxvals.data_ptr(), _ptr(coords), _ptr(idx)
    with torch.cuda.device(dev):
        L.Context.get(dev.index).run(L.OP_DECODE, a, torch.cuda.current_stream(dev).cuda_stream)
    decode_device._keep = (hm, center, scale)
    if want_coords:
        return preds, maxvals, coords, idx
    return preds, maxvals


def get_max_preds(batch_heatmaps):
    """heatmaps: numpy.ndarray([batch_size, num_joints, height, width]) -> (coords f32 [B,K,2], maxvals f32 [B,K,1])"""
    assert isinstance(batch_heatmaps, np.ndarray), "batch_heatmaps should be numpy.ndarray"
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).cuda()
    _, maxvals, coords, _ = decode_device(hm, None, None, post_process=False, want_coords=True)
    return coords.cpu().numpy(), maxvals.cpu().numpy()


def get_final_preds(config, batch_heatmaps, center, scale, use_dark=False):
    if use_dark is True:
        raise NotImplementedError("DARK decoding is disabled in the reference (use_dark=False default, "
                                  "lib/core/inference.py:51) and is out of scope")
    assert isinstance(batch_heatmaps, np.ndarray), "batch_heatmaps should be numpy.ndarray"
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).cuda()
    c = torch.from_numpy(np.ascontiguousarray(np.asarray(center), dtype=np.float32)).cuda()
    s = torch.from_numpy(np.ascontiguousarray(np.asarray(scale), dtype=np.float32)).cuda()
    preds, maxvals = decode_device(hm, c, s, post_process=bool(config.TEST.POST_PROCESS))
    return preds.cpu().numpy(), maxvals.cpu().numpy()
