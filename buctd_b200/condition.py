"""Condition rasteriser on the GPU: bottom-up keypoint proposals -> the blurred impulse maps the
top-down network is conditioned on.  Replaces ``get_stacked_condition`` /
``get_condition_image_colored`` / ``get_condition_image`` + ``generate_heatmap``,
/root/reference/lib/dataset/JointsDataset.py:457-543 (the reference does this per sample on the CPU
in DataLoader workers and ships 7.5 MB of float32 per crop over PCIe; here it is written straight
into the channels [3, 3+Cc) of the network input on the device)."""
from __future__ import annotations

import torch

from . import _lib as L

# /root/reference/lib/dataset/coco.py:54
COCO_KPT_COLORS = [[245, 59, 59], [249, 104, 25], [253, 183, 15], [233, 245, 41], [162, 252, 32],
                   [84, 247, 34], [31, 252, 57], [20, 246, 126], [5, 249, 206], [52, 215, 249],
                   [33, 136, 252], [11, 39, 248], [93, 46, 249], [156, 29, 244], [235, 49, 247],
                   [245, 47, 187], [253, 44, 117]]
MODES = {"stacked": 0, "colored": 1, "plain": 2}


def rasterize_condition(kpts: torch.Tensor, height: int, width: int, mode: str = "stacked", colors=None,
                        out: torch.Tensor | None = None, c_off: int = 0) -> torch.Tensor:
    """kpts: CUDA float32 [B,K,>=2] crop-pixel (x, y[, ...]).  Returns / fills float32
    ``out[B, C, H, W]`` channels [c_off, c_off+Cc), Cc = K (stacked) or 3 (colored / plain)."""
    if not kpts.is_cuda:
        raise L.BuctdError("rasterize_condition needs CUDA tensors; there is no CPU path")
    kp = kpts.contiguous().float()
    B, K, ks = kp.shape
    m = MODES[mode]
    cc = K if m == 0 else 3
    dev = kp.device
    if out is None:
        out = torch.empty(B, cc, height, width, dtype=torch.float32, device=dev)
        c_off = 0
    assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[0] == B and out.shape[2:] == (height, width)
    col = None
    if m == 1:
        col = torch.tensor(colors if colors is not None else [COCO_KPT_COLORS[i % 17] for i in range(K)],
                           dtype=torch.float32, device=dev).contiguous()
    scratch = torch.empty(max(B, 1), dtype=torch.int64, device=dev) if m != 0 else None
    a = L.RasterArgs()
    a.batch, a.k, a.kstride, a.h, a.w, a.mode = B, K, ks, height, width, m
    a.out_channels, a.c_off = out.shape[1], c_off
    a.kpts, a.colors, a.out = kp.data_ptr(), (col.data_ptr() if col is not None else None), out.data_ptr()
    a.scratch = scratch.data_ptr() if scratch is not None else None
    with torch.cuda.device(dev):
        L.Context.get(dev.index).run(L.OP_RASTER, a, torch.cuda.current_stream(dev).cuda_stream)
    rasterize_condition._keep = (kp, col, scratch)
    return out
