"""The optimizer of the reference's training step: ``get_optimizer`` builds ``torch.optim.Adam(params, lr=TRAIN.LR)``
(/root/reference/lib/utils/utils.py:268-272: weight decay and betas are NOT passed for adam).  Here all parameters are
re-homed into ONE flat float32 buffer (and their gradients into another), so one step is one CUDA kernel over
115.6 M elements instead of ~1800 small launches, and the data-parallel gradient reduction is ONE NCCL all-reduce of
the flat gradient buffer.

It is a ``torch.optim.Optimizer``: ``torch.optim.lr_scheduler.MultiStepLR(optimizer, ...)`` (tools/train.py:158-161)
drives ``param_groups[0]['lr']``, and ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's own format
(per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so the reference's checkpoints (tools/train.py:137,252) resume in
either direction."""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib as L
from .distributed import allreduce_mean_


class Adam(torch.optim.Optimizer):
    """Same update as torch.optim.Adam(params, lr) (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if weight_decay != 0.0:
            raise NotImplementedError("the reference passes no weight decay to Adam (lib/utils/utils.py:268-272)")
        super().__init__(params, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=0.0))
        self.params = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise L.BuctdError("buctd_b200.optim.Adam needs CUDA parameters; there is no CPU path")
        self.step_count = 0
        # 16-byte aligned slots so every tensor (and the whole buffer) can be processed with float4 accesses
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    def _check_homed(self):
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + 4 * self.flat.numel()
        for p, o in zip(self.params, self.offsets):
            if not (lo <= p.data_ptr() < hi):
                raise L.BuctdError("a parameter was moved out of the optimizer's flat buffer (model.cuda() / .to() after the optimizer was "
                                   "built): build buctd_b200.optim.Adam AFTER the model is on its device")
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)          # zero_grad(set_to_none=True) / user code dropped it

    def zero_grad(self, set_to_none: bool = False):
        """gradients live in the flat buffer: always zeroed in place (set_to_none would sever the views)"""
        self.flat_grad.zero_()
        self._check_homed()

    def allreduce_grads(self):
        """the ONE collective of the data-parallel training step: mean of the flat gradient buffer over ranks"""
        allreduce_mean_(self.flat_grad)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if any(abs(float(g["lr"]) - self.lr) > 0 for g in self.param_groups):
            raise NotImplementedError("one learning rate for all parameter groups (as the reference's get_optimizer)")
        self._check_homed()
        self.step_count += 1
        g0 = self.param_groups[0]
        a = L.AdamArgs()
        a.n, a.lr, a.beta1, a.beta2, a.eps, a.step = self.flat.numel(), self.lr, g0["betas"][0], g0["betas"][1], g0["eps"], self.step_count
        a.param, a.grad = self.flat.data_ptr(), self.flat_grad.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        dev = self.flat.device
        with torch.cuda.device(dev):
            L.Context.get(dev.index).call("buctd_adam_step", a, torch.cuda.current_stream(dev).cuda_stream)
        return loss

    # ---- checkpoints in torch.optim.Adam's format ---------------------------------------------------------------
    def state_dict(self):
        index = {}
        i = 0
        groups = []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                index[id(p)] = i
                ids.append(i)
                i += 1
            groups.append({**{k: v for k, v in g.items() if k != "params"}, "params": ids})
        state = {}
        for p, o in zip(self.params, self.offsets):
            if self.step_count == 0:
                continue
            n = p.numel()
            state[index[id(p)]] = {"step": torch.tensor(float(self.step_count)),
                                   "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                                   "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        order = [p for g in self.param_groups for p in g["params"]]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(order):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        for g, sg in zip(self.param_groups, groups):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
        off = {id(p): o for p, o in zip(self.params, self.offsets)}
        steps = set()
        with torch.no_grad():
            for i, p in zip(ids, order):
                st = state_dict["state"].get(i)
                if st is None or id(p) not in off:
                    continue
                o, n = off[id(p)], p.numel()
                self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise NotImplementedError("per-parameter step counts differ: one fused step count is kept")
        self.step_count = steps.pop() if steps else 0
