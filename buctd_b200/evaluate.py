"""Drop-ins for the two host-side helpers around the training step:

* ``generate_target``: ``JointsDataset.generate_target`` (/root/reference/lib/dataset/JointsDataset.py:397-453),
  batched: the Gaussian training heat-maps and their weights are written on the device instead of by the DataLoader
  workers (and copied host -> device every step);
* ``accuracy``: ``core.evaluate.accuracy`` (/root/reference/lib/core/evaluate.py:40-70), same return tuple; the
  reference copies the output heat-maps to the host for it every training step (lib/core/function.py:143).

Both run through the C ABI (buctd_generate_target, buctd_decode_heatmaps, buctd_pck_accuracy); there is no CPU path."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .inference import decode_device


def _stamp(sigma) -> np.ndarray:
    # the reference's own float32 expression (JointsDataset.py:428-433); computed with numpy on the host so the
    # table -- 361 floats for sigma 3 -- is the one the reference pastes
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    return np.ascontiguousarray(np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2)), dtype=np.float32)


def generate_target(joints, joints_vis, image_size, heatmap_size, sigma, joints_weight=None, device=None):
    """joints, joints_vis: [B,K,3] (array or tensor; float64 arithmetic like the reference).  image_size / heatmap_size:
    (w, h) as in cfg.MODEL.IMAGE_SIZE / HEATMAP_SIZE.  Returns CUDA tensors (target f32 [B,K,H,W], target_weight f32
    [B,K,1]).  ``joints_weight`` [K,1] = the dataset's joints_weight when LOSS.USE_DIFFERENT_JOINTS_WEIGHT is set."""
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise L.BuctdError("generate_target needs a CUDA device; there is no CPU path")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if int(sigma) != sigma:
        raise NotImplementedError("non-integer MODEL.SIGMA: the reference's int(mu - 3 sigma) arithmetic is only restated for integers")
    sigma = int(sigma)
    j = torch.as_tensor(np.asarray(joints) if not torch.is_tensor(joints) else joints).to(dev, torch.float64).contiguous()
    v = torch.as_tensor(np.asarray(joints_vis) if not torch.is_tensor(joints_vis) else joints_vis).to(dev, torch.float64).contiguous()
    assert j.dim() == 3 and j.shape[2] == 3 and v.shape == j.shape, "joints / joints_vis must be [B, K, 3]"
    B, K = j.shape[0], j.shape[1]
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    fs = np.asarray(image_size) / np.asarray(heatmap_size)
    stamp = torch.from_numpy(_stamp(sigma)).to(dev)
    jw = None
    if joints_weight is not None:
        jw = torch.as_tensor(np.asarray(joints_weight, dtype=np.float32)).reshape(K).to(dev).contiguous()
    target = torch.empty(B, K, H, W, dtype=torch.float32, device=dev)
    weight = torch.empty(B, K, 1, dtype=torch.float32, device=dev)
    a = L.GenTargetArgs()
    a.batch, a.k, a.h, a.w, a.tmp_size = B, K, H, W, 3 * sigma
    a.feat_stride_x, a.feat_stride_y = float(fs[0]), float(fs[1])
    a.joints, a.joints_vis, a.stamp = j.data_ptr(), v.data_ptr(), stamp.data_ptr()
    a.joints_weight = jw.data_ptr() if jw is not None else None
    a.target, a.target_weight = target.data_ptr(), weight.data_ptr()
    with torch.cuda.device(dev):
        L.Context.get(dev.index).call("buctd_generate_target", a, torch.cuda.current_stream(dev).cuda_stream)
    generate_target._keep = (j, v, stamp, jw)
    return target, weight


def accuracy_device(output: torch.Tensor, target: torch.Tensor, thr: float = 0.5):
    """CUDA float32 [B,K,H,W] heat-maps -> (acc float64 [K+1], cnt int32 [1], pred float32 [B,K,2]) CUDA tensors,
    nothing synchronised."""
    if not (output.is_cuda and target.is_cuda):
        raise L.BuctdError("accuracy_device needs CUDA tensors; there is no CPU path")
    B, K, H, W = output.shape
    dev = output.device
    _, _, pred, _ = decode_device(output, None, None, post_process=False, want_coords=True)
    _, _, tgt, _ = decode_device(target, None, None, post_process=False, want_coords=True)
    acc = torch.empty(K + 1, dtype=torch.float64, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    a = L.PckArgs()
    a.batch, a.k, a.h, a.w, a.thr = B, K, H, W, float(thr)
    a.pred, a.target, a.acc, a.cnt = pred.data_ptr(), tgt.data_ptr(), acc.data_ptr(), cnt.data_ptr()
    with torch.cuda.device(dev):
        L.Context.get(dev.index).call("buctd_pck_accuracy", a, torch.cuda.current_stream(dev).cuda_stream)
    accuracy_device._keep = (pred, tgt)
    return acc, cnt, pred


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """Reference signature: NumPy (or tensors) in, ``(acc ndarray [K+1], avg_acc, cnt, pred ndarray [B,K,2])`` out."""
    if hm_type != "gaussian":
        raise NotImplementedError("accuracy: only hm_type='gaussian' exists in the reference's callers")
    o = torch.as_tensor(np.ascontiguousarray(output, dtype=np.float32) if isinstance(output, np.ndarray) else output).cuda().float()
    t = torch.as_tensor(np.ascontiguousarray(target, dtype=np.float32) if isinstance(target, np.ndarray) else target).cuda().float()
    acc, cnt, pred = accuracy_device(o.contiguous(), t.contiguous(), thr)
    acc = acc.cpu().numpy()
    cnt = int(cnt.item())
    return acc, (float(acc[0]) if cnt != 0 else 0), cnt, pred.cpu().numpy()
