"""Drop-in for the flip-test helpers of the reference's ``utils.transforms`` (``flip_back``, the stacked-condition
branch of ``flip_hm``; /root/reference/lib/utils/transforms.py:16-58) and the device-side flip test of
``validate`` (/root/reference/lib/core/function.py:213-236).  The reference mirrors heat-maps on the host
(``.cpu().numpy()[..., ::-1]`` and back, one D2H + H2D round trip per batch); here both directions are one
CUDA kernel each on tensors that never leave the device."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib as L


def pairs_to_perm(n: int, pairs: Sequence[Sequence[int]]):
    """out[:, c] = in[:, perm[c]] equal to the reference's sequential pair swaps"""
    perm = list(range(n))
    for p0, p1 in pairs:
        perm[p0], perm[p1] = perm[p1], perm[p0]
    return perm


def _run(name, args, dev, keep):
    with torch.cuda.device(dev):
        L.Context.get(dev.index).call(name, args, torch.cuda.current_stream(dev).cuda_stream)
    _run._keep = keep


def flip_channels_device(x: torch.Tensor, perm: Sequence[int]) -> torch.Tensor:
    """out[b, c, y, x] = x[b, perm[c], y, W-1-x] (CUDA float32 NCHW)"""
    if not x.is_cuda:
        raise L.BuctdError("flip_channels_device needs a CUDA tensor; there is no CPU path")
    x = x.contiguous().float()
    B, Cn, H, W = x.shape
    if Cn > L.MAX_FLIP_CHANNELS:
        raise ValueError("at most %d channels" % L.MAX_FLIP_CHANNELS)
    out = torch.empty_like(x)
    a = L.FlipArgs()
    a.batch, a.channels, a.h, a.w = B, Cn, H, W
    for i, p in enumerate(perm):
        a.perm[i] = int(p)
    a.x, a.out = x.data_ptr(), out.data_ptr()
    _run("buctd_flip_channels", a, x.device, (x, out))
    return out


def flip_input_device(x: torch.Tensor, flip_pairs, stacked: bool) -> torch.Tensor:
    """``torch.cat((input[:, :3].flip(3), flip_hm(input[:, 3:], ...)), 1)`` (function.py:215) for a stacked condition
    (channel k of the condition = joint k: mirrored AND swapped by flip_pairs) or a condition that is only mirrored.
    A coloured condition is re-rasterised from the mirrored proposals in the reference (transforms.py:38-47):
    use ``condition.rasterize_condition`` on ``fliplr`` keypoints for that."""
    Cn = x.shape[1]
    perm = list(range(Cn))
    if stacked and Cn > 3:
        perm = [0, 1, 2] + [3 + p for p in pairs_to_perm(Cn - 3, flip_pairs)]
    return flip_channels_device(x, perm)


def flip_merge_device(output: torch.Tensor, output_flipped: torch.Tensor, flip_pairs, shift_heatmap: bool) -> torch.Tensor:
    """(output + shift(flip_back(output_flipped, flip_pairs))) * 0.5 in one kernel (function.py:226-236)"""
    if not (output.is_cuda and output_flipped.is_cuda):
        raise L.BuctdError("flip_merge_device needs CUDA tensors; there is no CPU path")
    o, f = output.contiguous().float(), output_flipped.contiguous().float()
    assert o.shape == f.shape and o.dim() == 4, "output_flipped should be [batch_size, num_joints, height, width]"
    B, K, H, W = o.shape
    out = torch.empty_like(o)
    a = L.FlipMergeArgs()
    a.batch, a.k, a.h, a.w, a.shift = B, K, H, W, int(bool(shift_heatmap))
    for i, p in enumerate(pairs_to_perm(K, flip_pairs)):
        a.perm[i] = int(p)
    a.hm, a.hm_flipped, a.out = o.data_ptr(), f.data_ptr(), out.data_ptr()
    _run("buctd_flip_merge", a, o.device, (o, f, out))
    return out


def flip_back(output_flipped: np.ndarray, matched_parts) -> np.ndarray:
    """reference signature: numpy.ndarray(batch_size, num_joints, height, width) in and out"""
    assert output_flipped.ndim == 4, "output_flipped should be [batch_size, num_joints, height, width]"
    t = torch.from_numpy(np.ascontiguousarray(output_flipped, dtype=np.float32)).cuda()
    return flip_channels_device(t, pairs_to_perm(t.shape[1], matched_parts)).cpu().numpy()
