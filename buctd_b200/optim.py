"""The optimizer of the reference's training step: ``get_optimizer`` builds ``torch.optim.Adam(params, lr=TRAIN.LR)``
(/root/reference/lib/utils/utils.py:268-272: weight decay and betas are NOT passed for adam).  Here all parameters are
re-homed into ONE flat float32 buffer (and their gradients into another), so one step is one CUDA kernel over
115.6 M elements instead of ~1800 small launches, and the data-parallel gradient reduction is ONE NCCL all-reduce of
the flat gradient buffer."""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib as L
from .distributed import allreduce_mean_


class Adam:
    """Same update as torch.optim.Adam(params, lr) (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise L.BuctdError("buctd_b200.optim.Adam needs CUDA parameters; there is no CPU path")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        # 16-byte aligned slots so every tensor (and the whole buffer) can be processed with float4 accesses
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.param_groups = [{"params": self.params, "lr": self.lr}]        # lr schedulers of the reference read/write this

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()

    def allreduce_grads(self):
        """the ONE collective of the data-parallel training step: mean of the flat gradient buffer over ranks"""
        allreduce_mean_(self.flat_grad)

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        self.lr = float(self.param_groups[0]["lr"])
        a = L.AdamArgs()
        a.n, a.lr, a.beta1, a.beta2, a.eps, a.step = self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.step_count
        a.param, a.grad = self.flat.data_ptr(), self.flat_grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        dev = self.flat.device
        with torch.cuda.device(dev):
            L.Context.get(dev.index).call("buctd_adam_step", a, torch.cuda.current_stream(dev).cuda_stream)
