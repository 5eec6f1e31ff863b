"""Drop-in for the reference's ``core.loss.JointsMSELoss`` (/root/reference/lib/core/loss.py:17-41): same constructor
and call signature, the value and its gradient with respect to ``output`` come from one fused CUDA pass
(the reference runs 17 separate nn.MSELoss reductions).  Closed form:
``0.5 / (K B HW) * sum (w (output - target))^2``."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L


class _JointsMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, target_weight, use_target_weight):
        if not output.is_cuda:
            raise L.BuctdError("JointsMSELoss needs CUDA tensors; there is no CPU path")
        o = output.contiguous().float()
        t = target.to(o.device).contiguous().float()
        B, K = o.shape[0], o.shape[1]
        hw = o[0, 0].numel()
        w = target_weight.to(o.device).reshape(B, K).contiguous().float() if use_target_weight else None
        loss = torch.empty((), dtype=torch.float32, device=o.device)
        grad = torch.empty_like(o)
        scratch = torch.empty(B * K, dtype=torch.float64, device=o.device)
        a = L.JointsMseArgs()
        a.batch, a.k, a.hw, a.use_target_weight = B, K, hw, int(bool(use_target_weight))
        a.output, a.target, a.target_weight = o.data_ptr(), t.data_ptr(), (w.data_ptr() if w is not None else None)
        a.loss, a.grad, a.scratch = loss.data_ptr(), grad.data_ptr(), scratch.data_ptr()
        with torch.cuda.device(o.device):
            L.Context.get(o.device.index).call("buctd_joints_mse", a, torch.cuda.current_stream(o.device).cuda_stream)
        ctx.save_for_backward(grad)
        ctx.keep = (o, t, w, scratch)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


class JointsMSELoss(nn.Module):
    def __init__(self, use_target_weight):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight):
        return _JointsMSE.apply(output, target, target_weight, bool(self.use_target_weight))
