"""The training step of the reference on the device (SURVEY.md section 8 row a-T):

    model.train(); output = model(input); loss = criterion(output, target, target_weight)
    optimizer.zero_grad(); loss.backward(); optimizer.step()           (/root/reference/lib/core/function.py:113-138)

``TrainStep(model, batch)`` records ONE step -- training-mode forward (BatchNorm with per-GPU batch statistics and
running-stat update, attention Dropout(0.1)), JointsMSELoss, the whole backward, Adam, and the re-pack of the updated
weights -- as a fixed sequence of C-ABI launches over a private workspace, and replays it as a CUDA graph.  Nothing here
computes through torch ops: torch owns the memory, the streams and (multi-GPU) the one NCCL all-reduce of the flat
gradient buffer.

Precision: 16-bit storage (bf16 by default, BASELINE.json configs[4]) for activations, activation gradients and the
packed copies of the weights; float32 master weights, BatchNorm statistics, parameter gradients and Adam state.
``loss_scale`` is a static scale on the loss gradient (needed for fp16 gradients; harmless for bf16), removed again when a
parameter gradient is written.

How the reference's autograd maps onto kernels:
  nn.Conv2d / nn.Linear   data gradient = tap-GEMM with transposed, tap-flipped packed weights (stride 2: on the zero-stuffed
                          gradient); weight gradient = buctd_gemm16 over the pixel rows (+ buctd_wgrad_reduce); bias = buctd_colsum
  nn.BatchNorm2d (+ReLU, +residual)   buctd_bn_stats / buctd_bn_apply forward, buctd_bn_bwd backward
  nn.Upsample + sum + ReLU (fuse)     buctd_hr_fuse / buctd_hr_fuse_bwd
  CoAM channel attention              buctd_coam_channel_att (dropout) ; gemm16 x4 + buctd_chan_softmax_bwd
  CoAM position attention             buctd_coam_position_att (dropout, log-sum-exp) ; rowdot + gemm16 x5 + buctd_pos_softmax_bwd
  JointsMSELoss, Adam                 buctd_joints_mse, buctd_adam_step
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .arch import BN_EPS, Conv, DAModule, HRNetArch, Linear
from .optim import Adam

BN_MOMENTUM = 0.1          # /root/reference/lib/models/pose_hrnet_coam.py:20


def _rup(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class _T:
    """an act16 tensor of the training workspace: [B][H+1][W+1][cs], zero pads; ``grad`` is allocated on first use"""
    __slots__ = ("ptr", "b", "h", "w", "c", "cs", "grad", "needs_grad", "name")

    def __init__(self, ptr, b, h, w, c, cs, needs_grad=True, name=""):
        self.ptr, self.b, self.h, self.w, self.c, self.cs = ptr, b, h, w, c, cs
        self.grad: Optional["_T"] = None
        self.needs_grad = needs_grad
        self.name = name

    @property
    def rows(self) -> int:
        return self.b * (self.h + 1) * (self.w + 1)

    @property
    def nbytes(self) -> int:
        return self.rows * self.cs * 2


def longest_final_range(poff: Dict[str, int], n: int, last: Dict[str, int], s: int):
    """(lo, hi): the longest contiguous element range of a flat buffer of ``n`` elements -- parameters ``poff`` = {name ->
    first element}, laid out back to back -- in which every parameter's last writer ``last[name]`` comes before op ``s``
    (a parameter without a writer never qualifies).  (0, 0) when there is none."""
    names = sorted(poff, key=poff.get)
    ends = [poff[k] for k in names[1:]] + [n]
    best, run_lo = (0, 0), None
    for k, end in zip(names, ends):
        if k in last and last[k] < s:
            if run_lo is None:
                run_lo = poff[k]
            if end - run_lo > best[1] - best[0]:
                best = (run_lo, end)
        else:
            run_lo = None
    return best


class TrainStep:
    """One recorded training step for a fixed batch size.  ``step(input, target, target_weight)`` returns the loss
    (a 0-d CUDA float32 tensor).  The model's parameters are re-homed into the optimizer's flat buffer (they stay
    ``nn.Parameter`` views: ``state_dict()`` / checkpoints work as before) and its BatchNorm running statistics are
    updated in place."""

    NLANES = 4        # side streams: one per branch of a HighResolutionModule

    def __init__(self, model, batch: int, lr: float = 1e-3, dtype: str = "bf16", loss_scale: float = 1.0, drop_p: float = 0.1,
                 seed: int = 0, use_target_weight: bool = True, optimizer: Optional[Adam] = None, graph: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise L.BuctdError("no CUDA device: buctd_b200 has no CPU path")
        arch: HRNetArch = model.arch
        if arch.use_prenet:
            raise NotImplementedError("training is built for the plain and the CoAM network (BASELINE.json configs[4]); preNet is inference only")
        if arch.coam and arch.att_channel_only:
            raise NotImplementedError("ATT_CHANNEL_ONLY training is not built")
        self.model, self.arch, self.B = model, arch, int(batch)
        self.dt = L.F16 if dtype == "fp16" else L.BF16
        self.tdt = torch.float16 if dtype == "fp16" else torch.bfloat16
        self.dtype = dtype
        self.loss_scale = float(loss_scale)
        self.inv_scale = 1.0 / self.loss_scale
        self.drop_p = float(drop_p) if arch.use_attention else 0.0
        self.base_seed = int(seed) & 0xFFFFFFFF
        self.use_target_weight = bool(use_target_weight)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise L.BuctdError("TrainStep needs the model on a CUDA device")
        self.ctx = L.Context.get(self.device.index)
        self.opt = optimizer or Adam(model.parameters(), lr=lr)
        self.poff: Dict[str, int] = {}
        off_of = {id(p): o for p, o in zip(self.opt.params, self.opt.offsets)}
        for name, p in model.named_parameters():
            self.poff[name] = off_of[id(p)]
        self.buffers = dict(model.named_buffers())
        self.steps_done = 0
        self.antialias = bool(getattr(model, "antialias", True))
        self.use_lanes = os.environ.get("BUCTD_TRAIN_LANES", "1") != "0"
        with torch.cuda.device(self.device):
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.NLANES)]
        # ---- two-pass build: measure the workspace, allocate, build with absolute pointers ----------
        self._build(dry=True)
        self.ws = torch.zeros(self._ws_bytes + (1 + self.NLANES) * (self._wg_need_final + 1024) + 4096, dtype=torch.uint8, device=self.device)
        self.w16 = torch.zeros(self._w16_elems + 64, dtype=self.tdt, device=self.device)
        self._build(dry=False)
        self.table = torch.tensor(self._table, dtype=torch.int64, device=self.device)
        units = []
        for e, (_, _, cout, cin, taps, _, _, _) in enumerate(self._table):
            assert taps <= 9
            for i in range((cout + 31) // 32):
                units.extend([e, i, j, 0] for j in range((cin + 31) // 32))
        self.units = torch.tensor(units, dtype=torch.int32, device=self.device)
        self.seed_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.seed_dev.zero_()
        self._repack()
        self._graph = None
        self._use_graph = (os.environ.get("BUCTD_TRAIN_GRAPH", "1") != "0") if graph is None else bool(graph)
        self._static_in = None
        self.on_grads_ready: Optional[Callable[[], None]] = None      # multi-GPU: the all-reduce of self.opt.flat_grad goes here
        # Bucketed overlap of that all-reduce with the backward (set_bucket_split): the backward is recorded in two parts;
        # after the first, the gradients of the flat buffer's range [lo, hi) are final and ``on_bucket_ready(lo, hi)`` may
        # start their (asynchronous) reduction while the second part runs; ``on_grads_ready`` then reduces the rest and waits.
        self.on_bucket_ready: Optional[Callable[[int, int], None]] = None
        self._split = None                  # (index into self.bwd, lo, hi): the early bucket is flat_grad[lo:hi]
        self._graph2 = None

    # ------------------------------------------------------------------------------------------------------
    # allocation helpers
    # ------------------------------------------------------------------------------------------------------
    def _take(self, nbytes: int) -> int:
        off = self._ws_off
        self._ws_off = _rup(off + nbytes, 1024)
        return self._ws_base + off

    def _act(self, h, w, c, cs=None, needs_grad=True, name="") -> _T:
        cs = cs or _rup(c, 8)
        return _T(self._take(self.B * (h + 1) * (w + 1) * cs * 2), self.B, h, w, c, cs, needs_grad, name)

    def _grad(self, t: _T):
        """(gradient tensor of t, accumulate?)  -- the first contribution writes, later ones add"""
        if t.grad is None:
            t.grad = _T(self._take(t.nbytes), t.b, t.h, t.w, t.c, t.cs, False, t.name + ".grad")
            return t.grad, 0
        return t.grad, 1

    def _pack(self, pname: str, cout: int, cin: int, taps: int, transposed: bool):
        """register a packed 16-bit copy of parameter `pname` ([cout][cin][taps] float32); returns (ptr, rows_alloc, cs)"""
        key = (pname, transposed)
        if key not in self._packed:
            rows = _rup(cin if transposed else cout, 16)
            cs = _rup(cout if transposed else cin, 8)
            off = self._w16_off
            self._w16_off = _rup(off + taps * rows * cs, 64)
            self._packed[key] = (off, rows, cs)
            self._table.append([self.poff[pname], off, cout, cin, taps, rows, cs, int(transposed)])
        off, rows, cs = self._packed[key]
        return self._w16_base + 2 * off, rows, cs

    def _pptr(self, pname: str) -> int:
        return self._flat_ptr + 4 * self.poff[pname]

    def _gptr(self, pname: str) -> int:
        return self._grad_ptr + 4 * self.poff[pname]

    def _call(self, lst: List, name: str, args):
        lst.append((name, args, self._cur_lane))

    # Lanes: the BasicBlocks of the branches of one HighResolutionModule are independent between two exchange / fuse steps, in
    # the forward and in the backward, and at 32 crops per GPU the low-resolution branches are rows of small, latency-bound
    # launches.  An op recorded while ``_cur_lane`` is i runs on side stream i; ops recorded with lane None run on the step's
    # stream and first join every active lane, and a lane forks from the step's stream when its first op arrives.  A backward
    # op inherits the lane its forward op was recorded on.  Reduction / split-K scratch is per lane.
    def _tape_add(self, fn):
        lane = self._cur_lane

        def wrapped():
            prev, self._cur_lane = self._cur_lane, lane
            try:
                fn()
            finally:
                self._cur_lane = prev
        self.tape.append(wrapped)

    def _lane_slot(self) -> int:
        return 0 if self._cur_lane is None else 1 + self._cur_lane

    # ------------------------------------------------------------------------------------------------------
    # op builders: each emits forward launches now and registers a backward builder on the tape
    # ------------------------------------------------------------------------------------------------------
    def _tapgemm_args(self, src: _T, w_ptr, n, n_alloc, w_cs, k, ntaps, out_ptr, out_cs, out_h, out_w, *, stride=1, bias=None,
                      residual: Optional[_T] = None, out_mode=L.OUT_ACT16, m_rows=None, a_cs=None, a_ptr=None):
        a = L.TapGemmArgs()
        a.dtype = self.dt
        a.mode = L.TG_PATCH if stride != 1 else L.TG_FLAT
        a.ntaps, a.stride = ntaps, stride
        a.batch, a.out_h, a.out_w = src.b, out_h, out_w
        a.in_h, a.in_w = src.h, src.w
        a.m_rows = m_rows if m_rows is not None else src.b * (out_h + 1) * (out_w + 1)
        a.k, a.a_cs = k, (a_cs if a_cs is not None else src.cs)
        a.n, a.n_alloc, a.w_cs = n, n_alloc, w_cs
        a.out_cs = out_cs
        a.res_cs = residual.cs if residual is not None else 0
        a.relu, a.out_mode = 0, out_mode
        a.a, a.w = (a_ptr if a_ptr is not None else src.ptr), w_ptr
        a.bias = bias
        a.residual = residual.ptr if residual is not None else None
        a.out = out_ptr
        return a

    def _wgrad(self, ops: List, dz: _T, x: _T, pname: str, cout: int, cin: int, ksize: int):
        """dW of a stride-1 convolution / linear layer whose output gradient dz and input x share their geometry"""
        taps = ksize * ksize
        rows = dz.rows
        g = L.Gemm16Args()
        g.dtype, g.a_mn_major, g.b_mn_major, g.m, g.n, g.k = self.dt, 1, 1, cout, cin, rows
        group = 3 if taps == 9 else 0         # the three dx taps of a kernel row share one A tile and one widened B stage
        bn = _rup(cin, 16) if cin <= (128 if group else 256) else (128 if group else 256)
        tiles = ((cout + 127) // 128) * ((cin + bn - 1) // bn) * (taps // 3 if group else taps)
        nsplit = max(1, min(148 // tiles if tiles <= 148 else 1, (rows + 255) // 256, 64))      # one wave of CTAs
        g.nbatch, g.nsplit, g.ntaps, g.b_rows, g.tap_group = 1, nsplit, taps, rows, group
        g.lda, g.ldb, g.ldd, g.out_f32, g.accumulate = dz.cs, x.cs, cin, 1, 0
        g.d_t_stride, g.d_s_stride, g.d_b_stride = cout * cin, taps * cout * cin, 0
        if taps == 9:
            for t in range(9):
                g.tap_off[t] = (t // 3 - 1) * (x.w + 1) + (t % 3 - 1)
        g.alpha = 1.0
        g.a, g.b, g.d = dz.ptr, x.ptr, self._wg_scratch[self._lane_slot()]
        self._wg_need = max(self._wg_need, nsplit * taps * cout * cin * 4)
        self._call(ops, "buctd_gemm16", g)
        r = L.WgradReduceArgs()
        r.n, r.k, r.ntaps, r.nsplit, r.accumulate, r.inv_scale = cout, cin, taps, nsplit, 0, self.inv_scale
        r.partial, r.grad = self._wg_scratch[self._lane_slot()], self._gptr(pname)
        self._call(ops, "buctd_wgrad_reduce", r)

    def _colsum(self, ops: List, ptr: int, c: int, cs: int, rows: int, pname: str):
        a = L.ColsumArgs()
        a.dtype, a.c, a.cs, a.accumulate, a.rows, a.inv_scale = self.dt, c, cs, 0, rows, self.inv_scale
        a.dy, a.dbias, a.scratch = ptr, self._gptr(pname), self._red_scratch[self._lane_slot()]
        self._call(ops, "buctd_colsum", a)

    def _conv_core(self, cv: Conv, x: _T, name=""):
        """z = conv(x) (no BatchNorm folded in); returns z and a function emitting the backward of the convolution"""
        oh, ow = (x.h, x.w) if cv.stride == 1 else (x.h // 2, x.w // 2)
        z = self._act(oh, ow, cv.cout, name=cv.w + ":z")
        taps = cv.k * cv.k
        wp, n_alloc, w_cs = self._pack(cv.w + ".weight", cv.cout, cv.cin, taps, False)
        bias = self._pptr(cv.w + ".bias") if cv.bias else None
        self._call(self.fwd, "buctd_tapgemm", self._tapgemm_args(x, wp, cv.cout, n_alloc, w_cs, cv.cin, taps, z.ptr, z.cs, oh, ow,
                                                                  stride=cv.stride, bias=bias))

        def backward(dz: _T):
            ops = self.bwd
            if cv.bias:
                self._colsum(ops, dz.ptr, cv.cout, dz.cs, dz.rows, cv.w + ".bias")
            g = dz
            if cv.stride == 2:                    # the stride-1 problem on the zero-stuffed gradient
                g = self._act(x.h, x.w, cv.cout, cs=dz.cs, needs_grad=False, name=cv.w + ":dz_up")
                s = L.ZeroStuffArgs()
                s.dtype, s.batch, s.h, s.w, s.cs = self.dt, self.B, oh, ow, dz.cs
                s.in_, s.out = dz.ptr, g.ptr
                self._call(ops, "buctd_zero_stuff", s)
            self._wgrad(ops, g, x, cv.w + ".weight", cv.cout, cv.cin, cv.k)
            if x.needs_grad:
                dx, acc = self._grad(x)
                wt, rows_t, cs_t = self._pack(cv.w + ".weight", cv.cout, cv.cin, taps, True)
                self._call(ops, "buctd_tapgemm", self._tapgemm_args(g, wt, cv.cin, rows_t, cs_t, cv.cout, taps, dx.ptr, dx.cs, x.h, x.w,
                                                                     residual=dx if acc else None))
        return z, backward

    def _bn_layer(self, bn: str, z: _T, residual: Optional[_T], relu: bool, conv_backward):
        """y = [ReLU](BN_train(z) [+ residual]); registers the backward of BN + the convolution behind it"""
        C_ = z.c
        y = self._act(z.h, z.w, C_, name=bn + ":y")
        mr = self._take(2 * C_ * 4)
        s = L.BnStatsArgs()
        s.dtype, s.batch, s.h, s.w, s.c, s.z_cs, s.eps, s.momentum = self.dt, self.B, z.h, z.w, C_, z.cs, BN_EPS, BN_MOMENTUM
        s.z, s.mean_rstd, s.scratch = z.ptr, mr, self._red_scratch[self._lane_slot()]
        s.running_mean, s.running_var = self.buffers[bn + ".running_mean"].data_ptr(), self.buffers[bn + ".running_var"].data_ptr()
        self._call(self.fwd, "buctd_bn_stats", s)
        a = L.BnApplyArgs()
        a.dtype, a.batch, a.h, a.w, a.c, a.z_cs, a.out_cs, a.relu = self.dt, self.B, z.h, z.w, C_, z.cs, y.cs, int(relu)
        a.res_cs = residual.cs if residual is not None else 0
        a.z, a.mean_rstd, a.gamma, a.beta = z.ptr, mr, self._pptr(bn + ".weight"), self._pptr(bn + ".bias")
        a.residual, a.out = (residual.ptr if residual is not None else None), y.ptr
        self._call(self.fwd, "buctd_bn_apply", a)
        self._nbt.append(bn + ".num_batches_tracked")

        def backward():
            if y.grad is None:
                return
            dz = self._act(z.h, z.w, C_, cs=z.cs, needs_grad=False, name=bn + ":dz")
            b = L.BnBwdArgs()
            b.dtype, b.batch, b.h, b.w, b.c = self.dt, self.B, z.h, z.w, C_
            b.dy_cs, b.y_cs, b.z_cs, b.dz_cs = y.grad.cs, y.cs, z.cs, dz.cs
            b.relu, b.inv_scale = int(relu), self.inv_scale
            b.dy, b.y, b.z, b.mean_rstd, b.gamma = y.grad.ptr, y.ptr, z.ptr, mr, self._pptr(bn + ".weight")
            b.dgamma, b.dbeta, b.scratch, b.dz = self._gptr(bn + ".weight"), self._gptr(bn + ".bias"), self._red_scratch[self._lane_slot()], dz.ptr
            if residual is not None and residual.needs_grad:
                dr, acc = self._grad(residual)
                b.dres, b.dres_cs, b.res_accumulate = dr.ptr, dr.cs, acc
            self._call(self.bwd, "buctd_bn_bwd", b)
            conv_backward(dz)
        self._tape_add(backward)
        return y

    def conv_bn(self, cv: Conv, x: _T, residual: Optional[_T] = None, relu: Optional[bool] = None) -> _T:
        z, cb = self._conv_core(cv, x)
        return self._bn_layer(cv.bn, z, residual, cv.relu if relu is None else relu, cb)

    def conv_bias(self, cv: Conv, x: _T) -> _T:
        """convolution with its own bias and no BatchNorm / activation (the CoAM convolutions)"""
        z, cb = self._conv_core(cv, x)

        def backward():
            if z.grad is not None:
                cb(z.grad)
        self._tape_add(backward)
        return z

    def linear(self, lin: Linear, x: _T, name="") -> _T:
        """nn.Linear over the channels of every pixel (the q / k / v / o projections)"""
        return self.conv_bias(Conv(lin.w, None, lin.fin, lin.fout, 1, 1, bias=True), x)

    def fuse(self, terms, h, w, c) -> _T:
        """out = ReLU(sum_j nearest_up(term_j)); terms: [(tensor, shift)]"""
        out = self._act(h, w, c, name="fuse")
        a = L.FuseArgs()
        a.dtype, a.batch, a.h, a.w, a.c, a.out_cs, a.nterms, a.relu = self.dt, self.B, h, w, c, out.cs, len(terms), 1
        for j, (t, sh) in enumerate(terms):
            a.term[j], a.term_cs[j], a.term_shift[j] = t.ptr, t.cs, sh
        a.out = out.ptr
        self._call(self.fwd, "buctd_hr_fuse", a)

        def backward():
            if out.grad is None:
                return
            b = L.FuseBwdArgs()
            b.dtype, b.batch, b.h, b.w, b.c, b.cs, b.relu = self.dt, self.B, h, w, c, out.cs, 1
            b.dout, b.out = out.grad.ptr, out.ptr
            n = 0
            for t, sh in terms:
                if not t.needs_grad:
                    continue
                g, acc = self._grad(t)
                b.dterm[n], b.term_cs[n], b.term_shift[n], b.term_accumulate[n] = g.ptr, g.cs, sh, acc
                n += 1
            b.nterms = n
            if n:
                self._call(self.bwd, "buctd_hr_fuse_bwd", b)
        self._tape_add(backward)
        return out

    # ---- CoAM --------------------------------------------------------------------------------------------
    def _gemm(self, ops, a_ptr, b_ptr, d_ptr, m, n, k, *, a_mn, b_mn, lda, ldb, ldd, nbatch=1, a_bs=0, b_bs=0, d_bs=0, out_f32=0, acc=0,
              alpha=1.0, b_rows=0):
        g = L.Gemm16Args()
        g.dtype, g.a_mn_major, g.b_mn_major, g.m, g.n, g.k = self.dt, int(a_mn), int(b_mn), m, n, k
        g.nbatch, g.nsplit, g.ntaps, g.b_rows = nbatch, 1, 1, b_rows
        g.lda, g.ldb, g.ldd, g.out_f32, g.accumulate = lda, ldb, ldd, int(out_f32), int(acc)
        g.a_batch_stride, g.b_batch_stride, g.d_b_stride = a_bs, b_bs, d_bs
        g.alpha = alpha
        g.a, g.b, g.d = a_ptr, b_ptr, d_ptr
        self._call(ops, "buctd_gemm16", g)

    def da_module(self, da: DAModule, x: _T, cond: _T, idx: int) -> _T:
        """DAModule.forward in training mode (pose_hrnet_coam.py:705-724): out = x + (p_out + c_out)"""
        B, C_, h, w = self.B, da.c, da.h, da.w
        hw, hws = h * w, _rup(h * w, 8)
        rows1 = (h + 1) * (w + 1)
        seed_c = (self.base_seed * 2654435761 + 2 * idx + 1) & 0xFFFFFFFF
        seed_p = (self.base_seed * 2654435761 + 2 * idx + 2) & 0xFFFFFFFF
        self.seeds[da.prefix + ":chan"], self.seeds[da.prefix + ":pos"] = seed_c, seed_p
        # ---------------- channel attention ----------------
        y = self.conv_bias(da.chan_cnn, x)
        yc = self.conv_bias(da.chan_cnn_cond, cond)
        planar = self._take(B * C_ * hws * 2)
        nch = (rows1 + L.CHAN_ATT_ROWS_PER_CHUNK - 1) // L.CHAN_ATT_ROWS_PER_CHUNK
        probs = self._take(B * nch * C_ * C_ * 4)
        ca = L.ChanAttArgs()
        ca.dtype, ca.batch, ca.h, ca.w, ca.c, ca.y_cs, ca.yc_cs, ca.out_stride = self.dt, B, h, w, C_, y.cs, yc.cs, hws
        ca.drop_p, ca.drop_seed, ca.seed_xor = self.drop_p, seed_c, self.seed_dev_ptr
        ca.y, ca.yc, ca.scratch, ca.out = y.ptr, yc.ptr, probs, planar
        self._call(self.fwd, "buctd_coam_channel_att", ca)
        lw = da.chan_fc_o.w
        c_out = self._take(B * C_ * hws * 2)
        wp, n_alloc, w_cs = self._pack(lw + ".weight", hw, hw, 1, False)
        src = _T(planar, 1, 0, 0, hw, hws, False)
        fa = self._tapgemm_args(src, wp, hw, n_alloc, w_cs, hw, 1, c_out, hws, 0, 0, bias=self._pptr(lw + ".bias"), out_mode=L.OUT_ROWS16,
                                m_rows=B * C_, a_cs=hws)
        fa.batch = 1
        self._call(self.fwd, "buctd_tapgemm", fa)
        # ---------------- position attention ----------------
        yp = self.conv_bias(da.pos_cnn, x)
        ycp = self.conv_bias(da.pos_cnn_cond, cond)
        q = self.linear(da.fc_q, ycp)
        k = self.linear(da.fc_k, yp)
        v = self.linear(da.fc_v, yp)
        o = self._act(h, w, C_, name=da.prefix + ":o")
        lse = self._take(B * rows1 * 4)
        pa = L.PosAttArgs()
        pa.dtype, pa.batch, pa.h, pa.w, pa.c = self.dt, B, h, w, C_
        pa.q_cs, pa.k_cs, pa.v_cs, pa.o_cs = q.cs, k.cs, v.cs, o.cs
        pa.drop_p, pa.drop_seed, pa.seed_xor = self.drop_p, seed_p, self.seed_dev_ptr
        pa.q, pa.k, pa.v, pa.o, pa.lse2 = q.ptr, k.ptr, v.ptr, o.ptr, lse

        def pos_backward():
            if o.grad is None:
                return
            ops = self.bwd
            do = o.grad
            ld = _rup(rows1, 8)
            delta = self._take(B * rows1 * 4)
            r = L.RowdotArgs()
            r.dtype, r.c, r.a_cs, r.b_cs, r.rows = self.dt, C_, do.cs, o.cs, B * rows1
            r.a, r.b, r.out = do.ptr, o.ptr, delta
            self._call(ops, "buctd_rowdot", r)
            pd = self._take(B * rows1 * ld * 2)
            ds = self._take(B * rows1 * ld * 2)
            if os.environ.get("BUCTD_POS_BWD_FUSED", "1") != "0":
                # S = Q K^T and dP = dO V^T in TMEM, soft-max backward in the epilogue: the float32 score matrices never exist
                f = L.PosBwdScoresArgs()
                f.dtype, f.batch, f.h, f.w, f.c, f.q_cs, f.k_cs, f.v_cs, f.do_cs, f.ld = self.dt, B, h, w, C_, q.cs, k.cs, v.cs, do.cs, ld
                f.drop_p, f.seed, f.seed_xor = self.drop_p, seed_p, self.seed_dev_ptr
                f.q, f.k, f.v, f.d_o, f.lse2, f.delta, f.pd, f.ds = q.ptr, k.ptr, v.ptr, do.ptr, lse, delta, pd, ds
                self._call(ops, "buctd_pos_att_bwd_scores", f)
            else:
                s32 = self._take(B * rows1 * ld * 4)
                dpd = self._take(B * rows1 * ld * 4)
                sc = math.log2(math.e) / math.sqrt(C_)
                self._gemm(ops, q.ptr, k.ptr, s32, rows1, rows1, C_, a_mn=0, b_mn=0, lda=q.cs, ldb=k.cs, ldd=ld, nbatch=B, a_bs=rows1 * q.cs,
                           b_bs=rows1 * k.cs, d_bs=rows1 * ld, out_f32=1, alpha=sc)
                self._gemm(ops, do.ptr, v.ptr, dpd, rows1, rows1, C_, a_mn=0, b_mn=0, lda=do.cs, ldb=v.cs, ldd=ld, nbatch=B, a_bs=rows1 * do.cs,
                           b_bs=rows1 * v.cs, d_bs=rows1 * ld, out_f32=1)
                e = L.PosSoftmaxBwdArgs()
                e.dtype, e.batch, e.h, e.w, e.ld, e.inv_sqrt_c, e.drop_p, e.seed = self.dt, B, h, w, ld, 1.0 / math.sqrt(C_), self.drop_p, seed_p
                e.s, e.dpd, e.lse2, e.delta, e.pd, e.ds, e.seed_xor = s32, dpd, lse, delta, pd, ds, self.seed_dev_ptr
                self._call(ops, "buctd_pos_softmax_bwd", e)
            dv, accv = self._grad(v)
            dk, acck = self._grad(k)
            dq, accq = self._grad(q)
            assert not (accv or acck or accq)
            self._gemm(ops, pd, do.ptr, dv.ptr, rows1, C_, rows1, a_mn=1, b_mn=1, lda=ld, ldb=do.cs, ldd=dv.cs, nbatch=B, a_bs=rows1 * ld,
                       b_bs=rows1 * do.cs, d_bs=rows1 * dv.cs)
            self._gemm(ops, ds, q.ptr, dk.ptr, rows1, C_, rows1, a_mn=1, b_mn=1, lda=ld, ldb=q.cs, ldd=dk.cs, nbatch=B, a_bs=rows1 * ld,
                       b_bs=rows1 * q.cs, d_bs=rows1 * dk.cs)
            self._gemm(ops, ds, k.ptr, dq.ptr, rows1, C_, rows1, a_mn=0, b_mn=1, lda=ld, ldb=k.cs, ldd=dq.cs, nbatch=B, a_bs=rows1 * ld,
                       b_bs=rows1 * k.cs, d_bs=rows1 * dq.cs)
        self._call(self.fwd, "buctd_coam_position_att", pa)
        self._tape_add(pos_backward)
        p_out = self.linear(da.fc_o, o)
        # ---------------- combine ----------------
        out = self._act(h, w, C_, name=da.prefix + ":out")
        cb = L.DaCombineArgs()
        cb.dtype, cb.batch, cb.h, cb.w, cb.c = self.dt, B, h, w, C_
        cb.x_cs, cb.p_cs, cb.out_cs, cb.channel_only, cb.c_stride = x.cs, p_out.cs, out.cs, 0, hws
        cb.x, cb.p_out, cb.c_out, cb.out = x.ptr, p_out.ptr, c_out, out.ptr
        self._call(self.fwd, "buctd_da_combine", cb)

        def combine_backward():
            if out.grad is None:
                return
            ops = self.bwd
            dout = out.grad
            # d x += dout
            dx, acc = self._grad(x)
            ad = L.Add16Args()
            ad.dtype, ad.cs_out, ad.cs_a, ad.c, ad.rows, ad.accumulate = self.dt, dx.cs, dout.cs, C_, dout.rows, acc
            ad.a, ad.out = dout.ptr, dx.ptr
            self._call(ops, "buctd_add16", ad)
            # d p_out = dout (p_out has no other consumer: its gradient IS dout)
            p_out.grad = dout
            # d c_out (planar) = dout^T
            dc = self._take(B * C_ * hws * 2)
            rl = L.RelayoutArgs()
            rl.dtype, rl.batch, rl.h, rl.w, rl.c, rl.cs, rl.p_stride, rl.accumulate = self.dt, B, h, w, C_, dout.cs, hws, 0
            rl.in_, rl.out = dout.ptr, dc
            self._call(ops, "buctd_act16_to_planar", rl)
            # ---- chan fc_o (HW x HW over the planar rows): bias, weight, data gradient
            self._colsum(ops, dc, hw, hws, B * C_, lw + ".bias")
            self._gemm(ops, dc, planar, self._gptr(lw + ".weight"), hw, hw, B * C_, a_mn=1, b_mn=1, lda=hws, ldb=hws, ldd=hw, out_f32=1,
                       alpha=self.inv_scale)
            dplanar = self._take(B * C_ * hws * 2)
            wt, rows_t, cs_t = self._pack(lw + ".weight", hw, hw, 1, True)
            srcd = _T(dc, 1, 0, 0, hw, hws, False)
            fb = self._tapgemm_args(srcd, wt, hw, rows_t, cs_t, hw, 1, dplanar, hws, 0, 0, out_mode=L.OUT_ROWS16, m_rows=B * C_, a_cs=hws)
            fb.batch = 1
            self._call(ops, "buctd_tapgemm", fb)
            # ---- channel attention core
            d_act = self._act(h, w, C_, cs=y.cs, needs_grad=False, name=da.prefix + ":d_att_out")
            rl2 = L.RelayoutArgs()
            rl2.dtype, rl2.batch, rl2.h, rl2.w, rl2.c, rl2.cs, rl2.p_stride, rl2.accumulate = self.dt, B, h, w, C_, d_act.cs, hws, 0
            rl2.in_, rl2.out = dplanar, d_act.ptr
            self._call(ops, "buctd_planar_to_act16", rl2)
            da_d = self._take(B * C_ * C_ * 4)
            self._gemm(ops, d_act.ptr, y.ptr, da_d, C_, C_, rows1, a_mn=1, b_mn=1, lda=d_act.cs, ldb=y.cs, ldd=C_, nbatch=B,
                       a_bs=rows1 * d_act.cs, b_bs=rows1 * y.cs, d_bs=C_ * C_, out_f32=1)
            ldc = _rup(C_, 8)
            ds16 = self._take(B * C_ * ldc * 2)
            pd16 = self._take(B * C_ * ldc * 2)
            sb = L.ChanSoftmaxBwdArgs()
            sb.dtype, sb.batch, sb.c, sb.ld, sb.prob_stride, sb.scale, sb.drop_p, sb.seed = self.dt, B, C_, ldc, nch * C_ * C_, 1.0 / math.sqrt(hw), self.drop_p, seed_c
            sb.probs, sb.da_d, sb.ds, sb.probs_d, sb.seed_xor = probs, da_d, ds16, pd16, self.seed_dev_ptr
            self._call(ops, "buctd_chan_softmax_bwd", sb)
            dy, accy = self._grad(y)
            dyc, accyc = self._grad(yc)
            assert not (accy or accyc)
            # dy[p][c2] = sum_c1 d_act[p][c1] A_d[c1][c2] + sum_c1 yc[p][c1] dS[c1][c2]
            self._gemm(ops, d_act.ptr, pd16, dy.ptr, rows1, C_, C_, a_mn=0, b_mn=1, lda=d_act.cs, ldb=ldc, ldd=dy.cs, nbatch=B,
                       a_bs=rows1 * d_act.cs, b_bs=C_ * ldc, d_bs=rows1 * dy.cs)
            self._gemm(ops, yc.ptr, ds16, dy.ptr, rows1, C_, C_, a_mn=0, b_mn=1, lda=yc.cs, ldb=ldc, ldd=dy.cs, nbatch=B,
                       a_bs=rows1 * yc.cs, b_bs=C_ * ldc, d_bs=rows1 * dy.cs, acc=1)
            # dyc[p][c1] = sum_c2 y[p][c2] dS[c1][c2]
            self._gemm(ops, y.ptr, ds16, dyc.ptr, rows1, C_, C_, a_mn=0, b_mn=0, lda=y.cs, ldb=ldc, ldd=dyc.cs, nbatch=B,
                       a_bs=rows1 * y.cs, b_bs=C_ * ldc, d_bs=rows1 * dyc.cs)
        # the combine backward must run BEFORE the backward of p_out's linear layer (it sets p_out.grad): p_out's layer was
        # taped earlier, so append this one last -- the tape is walked in reverse
        self._tape_add(combine_backward)
        return out

    # ------------------------------------------------------------------------------------------------------
    # the network
    # ------------------------------------------------------------------------------------------------------
    def _build(self, dry: bool):
        arch, B = self.arch, self.B
        self._ws_base = 0 if dry else _rup(self.ws.data_ptr(), 1024)
        self._w16_base = 0 if dry else _rup(self.w16.data_ptr(), 128)
        self._flat_ptr = self.opt.flat.data_ptr()
        self._grad_ptr = self.opt.flat_grad.data_ptr()
        self._ws_off, self._w16_off = 0, 0
        self._packed, self._table = {}, []
        self.fwd, self.bwd, self.tape, self._nbt = [], [], [], []
        self.seeds: Dict[str, int] = {}
        H, W = arch.img_h, arch.img_w
        H4, W4 = H // 4, W // 4
        self._cur_lane = None
        self._red_scratch = [self._take(8 * 2 * 2048 * (L.BN_BLOCKS + 1)) for _ in range(1 + self.NLANES)]
        self._wg_need = getattr(self, "_wg_need", 0) if not dry else 0
        self._wg_scratch = [self._take(self._wg_need_final if not dry else 0) for _ in range(1 + self.NLANES)]
        self.seed_dev = torch.zeros(1, dtype=torch.int32, device=self.device) if dry or not hasattr(self, "seed_dev") else self.seed_dev
        self.seed_dev_ptr = self.seed_dev.data_ptr()
        Cx = 3 + arch.in_cond
        self.in_channels = Cx
        self.x_in = self._take(B * Cx * H * W * 4)                     # nchw32 staging copy of the input
        self.tgt = self._take(B * arch.num_joints * H4 * W4 * 4)
        self.tw = self._take(B * arch.num_joints * 4)
        # ---- stem ----
        z1 = self._act(H // 2, W // 2, 64, name="conv1:z")
        sa = L.StemArgs()
        sa.dtype, sa.batch, sa.x_channels, sa.h, sa.w, sa.out_cs, sa.x_half, sa.no_relu = self.dt, B, Cx, H, W, z1.cs, 0, 1
        self.stem_bias0 = self._take(64 * 4)
        sa.x, sa.weight, sa.bias, sa.out = self.x_in, self._pptr("conv1.weight"), self.stem_bias0, z1.ptr
        self._call(self.fwd, "buctd_stem_conv", sa)

        def stem_backward(dz: _T):
            part = self._take(L.BN_BLOCKS * 64 * 27 * 4)
            s = L.StemWgradArgs()
            s.dtype, s.batch, s.x_channels, s.h, s.w, s.dz_cs, s.x_half = self.dt, B, Cx, H, W, dz.cs, 0
            s.x, s.dz, s.partial = self.x_in, dz.ptr, part
            self._call(self.bwd, "buctd_stem_wgrad", s)
            r = L.WgradReduceArgs()
            r.n, r.k, r.ntaps, r.nsplit, r.accumulate, r.inv_scale = 64, 27, 1, L.BN_BLOCKS, 0, self.inv_scale
            r.partial, r.grad = part, self._gptr("conv1.weight")
            self._call(self.bwd, "buctd_wgrad_reduce", r)
        y1 = self._bn_layer("bn1", z1, None, True, stem_backward)
        cur = self.conv_bn(arch.stem[1], y1)
        # ---- layer1 ----
        for blk in arch.layer1:
            t = self.conv_bn(blk.convs[0], cur)
            t = self.conv_bn(blk.convs[1], t)
            res = cur
            if blk.downsample is not None:
                res = self.conv_bn(blk.downsample, cur)
            cur = self.conv_bn(blk.convs[2], t, residual=res, relu=True)
        # ---- condition maps (no gradient) ----
        cond: Dict[int, _T] = {}
        if arch.use_attention:
            nbr = max(len(v) for v in arch.coam.values())
            for i in range(nbr):
                cond[i] = self._act(H >> (2 + i), W >> (2 + i), arch.d_cond, cs=_rup(arch.d_cond, 16), needs_grad=False, name="cond%d" % i)
                ra = L.ResizeArgs()
                ra.dtype, ra.batch, ra.x_channels, ra.c_off, ra.cc = self.dt, B, Cx, 3, arch.d_cond
                ra.h, ra.w, ra.oh, ra.ow, ra.out_cs, ra.antialias = H, W, cond[i].h, cond[i].w, cond[i].cs, int(self.antialias)
                ra.x, ra.out = self.x_in, cond[i].ptr
                self._call(self.fwd, "buctd_cond_resize", ra)
        # ---- stages ----
        ys = [cur]
        da_idx = 0
        for si, (trans, mods) in enumerate(zip(arch.transitions, arch.stages)):
            chans = arch.stage_channels[si]
            nb = len(chans)
            xs = []
            for i in range(nb):
                t = trans[i]
                if t is None:
                    xs.append(ys[i])
                elif i < len(ys):
                    xs.append(self.conv_bn(t[0], ys[i]))
                else:
                    src = ys[-1]
                    for cv in t:
                        src = self.conv_bn(cv, src)
                    xs.append(src)
            if si in arch.coam and arch.use_attention:
                new = []
                for i, (da, x) in enumerate(zip(arch.coam[si], xs)):
                    new.append(self.da_module(da, x, cond[i], da_idx))
                    da_idx += 1
                xs = new
            for m in mods:
                for i in range(nb):
                    self._cur_lane = i if (self.use_lanes and nb > 1) else None
                    for blk in m.branches[i]:
                        t = self.conv_bn(blk.convs[0], xs[i])
                        xs[i] = self.conv_bn(blk.convs[1], t, residual=xs[i], relu=True)
                self._cur_lane = None
                if not m.fuse:
                    continue
                outs = []
                for i, terms in enumerate(m.fuse):
                    ft = []
                    for t in terms:
                        if t.kind == "id":
                            ft.append((xs[t.src], 0))
                        elif t.kind == "up":
                            ft.append((self.conv_bn(t.convs[0], xs[t.src]), t.src - i))
                        else:
                            src = xs[t.src]
                            for cv in t.convs:
                                src = self.conv_bn(cv, src)
                            ft.append((src, 0))
                    outs.append(self.fuse(ft, H4 >> i, W4 >> i, chans[i]))
                xs = outs
            ys = xs
        if 3 in arch.coam and arch.use_attention:
            ys = [self.da_module(arch.coam[3][0], ys[0], cond[0], da_idx)] + list(ys[1:])
        # ---- head + loss ----
        fs = ys[0]
        K = arch.num_joints
        fl = arch.final
        self.hm = self._take(B * K * H4 * W4 * 4)
        wp, n_alloc, w_cs = self._pack("final_layer.weight", K, fl.cin, fl.k * fl.k, False)
        self._call(self.fwd, "buctd_tapgemm", self._tapgemm_args(fs, wp, K, n_alloc, w_cs, fl.cin, fl.k * fl.k, self.hm, 0, fs.h, fs.w,
                                                                  bias=self._pptr("final_layer.bias"), out_mode=L.OUT_NCHW32))
        self.loss_t = torch.zeros((), dtype=torch.float32, device=self.device) if dry or not hasattr(self, "loss_t") else self.loss_t
        dhm32 = self._take(B * K * H4 * W4 * 4)
        mse_scratch = self._take(B * K * 8)
        ms = L.JointsMseArgs()
        ms.batch, ms.k, ms.hw, ms.use_target_weight, ms.grad_scale = B, K, H4 * W4, int(self.use_target_weight), self.loss_scale
        ms.output, ms.target, ms.target_weight = self.hm, self.tgt, self.tw
        ms.loss, ms.grad, ms.scratch = self.loss_t.data_ptr(), dhm32, mse_scratch
        self._call(self.fwd, "buctd_joints_mse", ms)
        # backward of the head: dhm (nchw32) -> act16, bias / weight / data gradient
        dhm = self._act(H4, W4, K, cs=_rup(K, 8), needs_grad=False, name="dhm")
        self.bwd.append(("nchw32_to_act16", (dhm32, B, K, 0, K, H4, W4, dhm.cs, dhm.ptr), None))
        self._colsum(self.bwd, dhm.ptr, K, dhm.cs, dhm.rows, "final_layer.bias")
        self._wgrad(self.bwd, dhm, fs, "final_layer.weight", K, fl.cin, fl.k)
        dfs, acc = self._grad(fs)
        wt, rows_t, cs_t = self._pack("final_layer.weight", K, fl.cin, fl.k * fl.k, True)
        self._call(self.bwd, "buctd_tapgemm", self._tapgemm_args(dhm, wt, fl.cin, rows_t, cs_t, K, fl.k * fl.k, dfs.ptr, dfs.cs, fs.h, fs.w,
                                                                  residual=dfs if acc else None))
        for fn in reversed(self.tape):
            fn()
        self._ws_bytes = self._ws_off
        self._w16_elems = self._w16_off
        self._wg_need_final = _rup(self._wg_need, 1024)
        self.heatmap_shape = (B, K, H4, W4)

    # ------------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------------
    def _run_ops(self, ops, stream):
        """``stream``: the torch stream of the step.  Ops with a lane run on that lane's side stream (forked from / joined into
        ``stream`` with events: inside a CUDA-graph capture these become the graph's dependency edges)."""
        lib, h = self.ctx.lib, self.ctx.handle
        main = stream.cuda_stream
        active = []

        def join():
            for l in active:
                ev = torch.cuda.Event()
                ev.record(self._side[l])
                stream.wait_event(ev)
            del active[:]

        for name, args, lane in ops:
            if lane is None:
                if active:
                    join()
                st = main
            else:
                if lane not in active:
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    self._side[lane].wait_event(ev)
                    active.append(lane)
                st = self._side[lane].cuda_stream
            if name == "nchw32_to_act16":
                self.ctx.nchw32_to_act16(self.dt, *args, st)
            else:
                L.check(getattr(lib, name)(h, C.byref(args), C.c_void_p(st)), name)
        join()

    def _repack(self):
        a = L.RepackArgs()
        a.dtype, a.n_entries, a.n_units = self.dt, self.table.shape[0], self.units.shape[0]
        a.param, a.w16, a.table, a.units = self.opt.flat.data_ptr(), self._w16_base, self.table.data_ptr(), self.units.data_ptr()
        with torch.cuda.device(self.device):
            self.ctx.call("buctd_repack_weights", a, torch.cuda.current_stream(self.device).cuda_stream)

    def _copy_in(self, input, target, target_weight):
        B, K = self.B, self.arch.num_joints
        if input.shape[0] != B or input.shape[1] < self.in_channels:
            raise ValueError("TrainStep was recorded for batch %d and %d input channels, got %s" % (B, self.in_channels, tuple(input.shape)))
        views = self._io_views
        views[0].copy_(input[:, :self.in_channels], non_blocking=True)
        views[1].copy_(target.reshape(views[1].shape), non_blocking=True)
        views[2].copy_(target_weight.reshape(B, K), non_blocking=True)

    @property
    def _io_views(self):
        if getattr(self, "_views", None) is None:
            B, K = self.B, self.arch.num_joints
            H, W = self.arch.img_h, self.arch.img_w
            base = self.ws.data_ptr()

            def view(ptr, shape):
                n = int(np.prod(shape))
                off = ptr - base
                return self.ws[off:off + 4 * n].view(torch.float32).view(*shape)
            self._views = (view(self.x_in, (B, self.in_channels, H, W)), view(self.tgt, (B, K, H // 4, W // 4)), view(self.tw, (B, K)),
                           view(self.hm, self.heatmap_shape))
        return self._views

    def _fwd_bwd(self, stream, part=None):
        """stream: the (current) torch stream of the step.  part: None = the whole step; 0 = forward + the backward up to the
        bucket split; 1 = the rest of the backward"""
        if part in (None, 0):
            self.seed_dev.copy_(self.seed_host, non_blocking=True)
            self._run_ops(self.fwd, stream)
        if part is None:
            self._run_ops(self.bwd, stream)
        elif part == 0:
            self._run_ops(self.bwd[:self._split[0]], stream)
        else:
            self._run_ops(self.bwd[self._split[0]:], stream)

    def grad_last_writes(self) -> Dict[str, int]:
        """{parameter name -> index of the LAST backward op that writes its gradient}: every pointer-valued field of every
        recorded backward op that points into the flat gradient buffer (found by value, not by bookkeeping)"""
        lo = self._grad_ptr
        hi = lo + 4 * self.opt.flat_grad.numel()
        names = sorted(self.poff, key=self.poff.get)
        starts = [self.poff[n] for n in names]
        import bisect
        last: Dict[str, int] = {}
        for i, (_, args, _) in enumerate(self.bwd):
            vals = args if isinstance(args, tuple) else [getattr(args, f[0]) for f in args._fields_]
            for v in vals:
                if isinstance(v, int) and lo <= v < hi:
                    last[names[bisect.bisect_right(starts, (v - lo) // 4) - 1]] = i
        return last

    def set_bucket_split(self, frac: float = 0.5):
        """Cut the recorded backward after ``frac`` of its ops and return ``(lo, hi)``: the longest contiguous range of the
        flat gradient buffer whose parameters all have their FINAL gradient once the first part has run (the backward walks
        the network from the head to the stem and the buffer is in module-registration order: the range is the late stages).
        Call before the first step."""
        if self._graph is not None:
            raise RuntimeError("set_bucket_split() must be called before the step is recorded as a graph")
        s = max(1, min(len(self.bwd) - 1, int(len(self.bwd) * frac)))
        last = self.grad_last_writes()
        best = longest_final_range(self.poff, self.opt.flat_grad.numel(), last, s)
        self._split = (s, best[0], best[1])
        return best

    def forward_backward(self, input, target, target_weight) -> torch.Tensor:
        """forward + loss + backward: fills ``self.opt.flat_grad`` (every parameter's ``.grad`` is a view of it)"""
        word = self.step_word(self.steps_done)
        self.seed_host[0] = word - (1 << 32) if word >= (1 << 31) else word
        with torch.cuda.device(self.device):
            self._copy_in(input, target, target_weight)
            stream = torch.cuda.current_stream(self.device)
            if self._split is not None:
                self._fwd_bwd_split(stream)
            elif not self._use_graph:
                self._fwd_bwd(stream)
            else:
                if self._graph is None:
                    self._fwd_bwd(stream)                          # eager warm-up run (also validates every launch)
                    stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._fwd_bwd(torch.cuda.current_stream(self.device))
                    self._graph = g
                else:
                    self._graph.replay()
            torch._foreach_add_([self.buffers[n] for n in self._nbt], 1)
        return self.loss_t

    def _fwd_bwd_split(self, stream):
        """the step in two parts with ``on_bucket_ready(lo)`` between them (two CUDA graphs when graphs are on)"""
        def hook():
            if self.on_bucket_ready is not None:
                self.on_bucket_ready(self._split[1], self._split[2])
        if not self._use_graph or self._graph is None:
            self._fwd_bwd(stream, 0)                               # eager (first step: also validates every launch)
            hook()
            self._fwd_bwd(stream, 1)
            if self._use_graph:
                stream.synchronize()
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    self._fwd_bwd(torch.cuda.current_stream(self.device), 0)
                with torch.cuda.graph(g2):
                    self._fwd_bwd(torch.cuda.current_stream(self.device), 1)
                self._graph, self._graph2 = g1, g2
        else:
            self._graph.replay()
            hook()
            self._graph2.replay()

    def step(self, input, target, target_weight) -> torch.Tensor:
        loss = self.forward_backward(input, target, target_weight)
        if self.on_grads_ready is not None:
            self.on_grads_ready()
        self.opt.step()
        self._repack()
        self.model.invalidate()            # the inference path re-folds / re-packs from the updated parameters on next use
        self.steps_done += 1
        return loss

    def step_word(self, step: int) -> int:
        """the per-step word XOR-ed into every dropout seed (a device word, so a replayed graph sees a new mask each step)"""
        return (step * 0x9E3779B1) & 0xFFFFFFFF if self.drop_p > 0 else 0

    def effective_seeds(self, step: int) -> Dict[str, int]:
        """{"<DAModule prefix>:chan" | ":pos" -> the uint32 dropout seed in force at `step`} (what the test oracle needs)"""
        return {k: v ^ self.step_word(step) for k, v in self.seeds.items()}

    @property
    def heatmaps(self) -> torch.Tensor:
        """the training-mode output of the last step ([B, K, H/4, W/4] float32 view of the workspace)"""
        return self._io_views[3]

    def launches_per_step(self) -> int:
        return len(self.fwd) + len(self.bwd) + 2
