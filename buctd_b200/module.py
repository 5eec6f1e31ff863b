"""Host mirror of the reference's ``PoseHighResolutionNet`` (both variants) on top of the C ABI.

Drop-in surface: ``get_pose_net(cfg, is_train, **kwargs) -> nn.Module`` whose
``forward(x[B, 3+Cc, H, W] float32) -> [B, K, H/4, W/4] float32`` on CUDA, whose ``state_dict()``
has exactly the reference's keys (``load_state_dict(strict=True)`` of a reference checkpoint works)
and which raises the same exceptions (``"condition is empty"``, branch-count ``ValueError``).
Reference: /root/reference/lib/models/pose_hrnet_coam.py:277-618, /root/reference/lib/models/pose_hrnet.py:276-495.

The module keeps the reference-named parameters as the source of truth.  On first use (and after
every ``load_state_dict``) they are packed once for the kernels: eval-mode BatchNorm folded into the
convolution in fp32, cast to the 16-bit compute type, laid out ``[tap][cout][cin]``.  One forward is
a recorded plan of C-ABI launches (tap-GEMMs on tcgen05, CUDA-core kernels for the HBM-bound
steps) replayed with a single call; nothing in this file computes on the CPU or through torch ops.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from .arch import BN_EPS, Conv, DAModule, HRNetArch, Linear

logger = logging.getLogger(__name__)


def _rup(v: int, m: int) -> int:
    return (v + m - 1) // m * m


@dataclass
class Act:
    """An act16 buffer inside the workspace: [B][H+1][W+1][cs] 16-bit."""
    ptr: int
    b: int
    h: int
    w: int
    c: int
    cs: int

    @property
    def rows(self) -> int:
        return self.b * (self.h + 1) * (self.w + 1)


class _Bump:
    def __init__(self):
        self.off = 0
        self.items = []

    def take(self, nbytes: int) -> int:
        off = self.off
        self.off = _rup(off + nbytes, 1024)
        return off


class PoseHighResolutionNet(nn.Module):
    def __init__(self, cfg, dtype: Optional[str] = None, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.arch = HRNetArch(cfg)
        dtype = dtype or os.environ.get("BUCTD_DTYPE", "fp16")
        if dtype not in ("fp16", "bf16"):
            raise ValueError("dtype must be 'fp16' or 'bf16', got %r" % (dtype,))
        self.compute_dtype = dtype
        self._dt = L.F16 if dtype == "fp16" else L.BF16
        self._torch_dt = torch.float16 if dtype == "fp16" else torch.bfloat16
        # TF.resize of the condition (lib/models/pose_hrnet_coam.py:755): torchvision >= 0.17 antialiases tensors by default,
        # torchvision 0.9 (the authors' training environment) did not -- checkpoints trained there want antialias=False
        # (kwarg, or cfg.MODEL.COND_RESIZE_ANTIALIAS)
        aa = kwargs.get("antialias")
        if aa is None:
            try:
                aa = cfg.MODEL.COND_RESIZE_ANTIALIAS
            except (AttributeError, KeyError):
                aa = True
        self.antialias = bool(aa)
        for name, shape, kind in self.arch.param_specs():
            self._register(name, shape, kind)
        self._packed = None          # dict of device tensors
        self._plans: Dict[tuple, dict] = {}
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    # ---- parameter container ------------------------------------------------------------------
    def _register(self, name: str, shape, kind: str):
        parts = name.split(".")
        mod = self
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        leaf = parts[-1]
        if kind in ("bn_rm", "bn_rv"):
            mod.register_buffer(leaf, torch.zeros(shape) if kind == "bn_rm" else torch.ones(shape))
        elif kind == "bn_nbt":
            mod.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
        else:
            init = torch.ones(shape) if kind == "bn_w" else torch.zeros(shape)
            mod.register_parameter(leaf, nn.Parameter(init))

    def init_weights(self, pretrained: str = ""):
        """Same contract as the reference (pose_hrnet_coam.py:574-609): conv / linear N(0, 1e-3),
        biases 0, BN 1/0; then a partial load filtered by MODEL.EXTRA.PRETRAINED_LAYERS."""
        logger.info("=> init weights from normal distribution")
        with torch.no_grad():
            for name, shape, kind in self.arch.param_specs():
                t = self._get(name)
                if kind in ("conv_w", "lin_w"):
                    t.normal_(std=0.001)
                elif kind in ("conv_b", "lin_b", "bn_b"):
                    t.zero_()
                elif kind == "bn_w":
                    t.fill_(1.0)
        if os.path.isfile(pretrained):
            sd = torch.load(pretrained, map_location="cpu")
            layers = list(self.cfg["MODEL"]["EXTRA"]["PRETRAINED_LAYERS"])
            keep = {k: v for k, v in sd.items() if k.split(".")[0] in layers or layers[0] == "*"}
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            logger.error("=> please download pre-trained models first!")
            raise ValueError("{} is not exist!".format(pretrained))
        self.invalidate()

    def _get(self, name: str) -> torch.Tensor:
        mod = self
        for p in name.split("."):
            mod = getattr(mod, p)
        return mod

    def invalidate(self):
        """Drop packed weights and plans (call after changing parameters in place)."""
        self._packed = None
        self._plans = {}

    def _apply(self, fn, *a, **k):   # .cuda() / .to(): packed copies follow lazily
        self.invalidate()
        return super()._apply(fn, *a, **k)

    # ---- weight packing -------------------------------------------------------------------------
    def _fold(self, conv: Conv):
        """BN(conv(x)) = conv'(x) + b' with w' = w * g/sqrt(v+eps), b' = beta + (bias - mean) * g/sqrt(v+eps)
        (eval-mode nn.BatchNorm2d, eps 1e-5), all in float64 before the cast."""
        w = self._get(conv.w + ".weight").detach().cpu().double()
        b = self._get(conv.w + ".bias").detach().cpu().double() if conv.bias else torch.zeros(conv.cout, dtype=torch.float64)
        if conv.bn:
            g = self._get(conv.bn + ".weight").detach().cpu().double()
            beta = self._get(conv.bn + ".bias").detach().cpu().double()
            mean = self._get(conv.bn + ".running_mean").detach().cpu().double()
            var = self._get(conv.bn + ".running_var").detach().cpu().double()
            s = g / torch.sqrt(var + BN_EPS)
            w = w * s.view(-1, 1, 1, 1)
            b = beta + (b - mean) * s
        return w, b

    def _pack(self, device):
        arch = self.arch
        wchunks, bchunks = [], []
        woff = boff = 0
        index: Dict[str, dict] = {}

        def add_w(key, w4, bias, ):
            nonlocal woff, boff
            cout, cin, kh, kw = w4.shape
            # rows per tap: a multiple of 16 (zero rows beyond cout).  BUCTD_PAD48=1 pads 48 output channels to a 64-wide N
            # tile: measured neutral (37.4 vs 37.7 us per conv), the operand fetch is per 128-byte row, DESIGN.md 3.1
            n_alloc, k_cs = (64 if cout == 48 and os.environ.get("BUCTD_PAD48", "0") == "1" else _rup(cout, 16)), _rup(cin, 8)
            p = torch.zeros(kh * kw, n_alloc, k_cs, dtype=torch.float32)
            p[:, :cout, :cin] = w4.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).float()
            p = p.to(self._torch_dt).reshape(-1)
            index[key] = {"w_off": woff, "b_off": boff, "n": cout, "k": cin, "n_alloc": n_alloc, "w_cs": k_cs,
                          "ntaps": kh * kw}
            wchunks.append(p)
            pad = (-p.numel()) % 64            # keep every matrix 128-byte aligned
            if pad:
                wchunks.append(torch.zeros(pad, dtype=self._torch_dt))
            woff += p.numel() + pad
            bchunks.append(bias.float())
            padb = (-bias.numel()) % 4
            if padb:
                bchunks.append(torch.zeros(padb))
            boff += bias.numel() + padb

        f32: Dict[str, torch.Tensor] = {}
        for conv in arch.convs():
            if conv.w == "conv1" or conv.w.startswith("rgb_preNet") or conv.w.startswith("cond_preNet"):
                continue
            w, b = self._fold(conv)
            add_w(conv.w, w, b)
        # k|v projections share their input: one GEMM with N = 2C
        for s in arch.coam:
            for da in arch.coam[s]:
                for lin in ([da.fc_q, da.fc_o] if not arch.att_channel_only else []) + [da.chan_fc_o]:
                    w = self._get(lin.w + ".weight").detach().cpu().double()
                    if lin is da.chan_fc_o:
                        # the HW x HW projection reads the channel-attention output in PADDED pixel order (buctd_coam_channel_att,
                        # out_padded): its weight matrix gets zero columns at the pad positions, K = (H+1)(W+1)
                        rows1 = (da.h + 1) * (da.w + 1)
                        valid = (torch.arange(da.h).view(-1, 1) * (da.w + 1) + torch.arange(da.w).view(1, -1)).reshape(-1)
                        wp = torch.zeros(lin.fout, rows1, dtype=torch.float64)
                        wp[:, valid] = w
                        add_w(lin.w, wp.view(lin.fout, rows1, 1, 1), self._get(lin.w + ".bias").detach().cpu().double())
                        continue
                    add_w(lin.w, w.view(lin.fout, lin.fin, 1, 1), self._get(lin.w + ".bias").detach().cpu().double())
                if not arch.att_channel_only:
                    wk = self._get(da.fc_k.w + ".weight").detach().cpu().double()
                    wv = self._get(da.fc_v.w + ".weight").detach().cpu().double()
                    bk = self._get(da.fc_k.w + ".bias").detach().cpu().double()
                    bv = self._get(da.fc_v.w + ".bias").detach().cpu().double()
                    add_w(da.prefix + ":kv", torch.cat([wk, wv], 0).view(2 * da.c, da.c, 1, 1), torch.cat([bk, bv], 0))
                    # the two 3x3 convolutions of a DAModule that read x (channel / position branch) likewise: N = 2C
                    w1, b1 = self._fold(da.chan_cnn)
                    w2, b2 = self._fold(da.pos_cnn)
                    add_w(da.prefix + ":cnn2", torch.cat([w1, w2], 0), torch.cat([b1, b2], 0))
        # stem conv1 runs on CUDA cores in fp32: [64][27]
        w, b = self._fold(arch.stem[0])
        f32["conv1.w"] = w.reshape(64, 27).float().contiguous()
        f32["conv1.b"] = b.float().contiguous()
        if arch.use_prenet:
            f32.update(self._compose_prenet())
        packed = {"w16": torch.cat(wchunks).to(device), "b32": torch.cat(bchunks).to(device), "index": index,
                  "f32": {k: v.to(device) for k, v in f32.items()}, "device": device}
        return packed

    def _compose_prenet(self):
        """rgb_preNet = conv3x3(3->64,bias)+BN ; conv7x7(64->3,bias)+BN with no activation between them
        (pose_hrnet.py:431-442): composed in float64 into 49 border variants of one 9x9 3->3 kernel
        (see include/buctd_b200.h, buctd_prenet_args)."""
        a = self.arch
        w1, b1 = self._fold(a.prenet_rgb[0])      # [64,3,3,3], [64]
        w2, b2 = self._fold(a.prenet_rgb[1])      # [3,64,7,7], [3]
        wc, bc = self._fold(a.prenet_cond[0])     # [3,3,7,7], [3]
        # contribution of 7x7 tap (ky,kx): w2[:, :, ky, kx] @ w1 placed at offset (ky, kx) of the 9x9 kernel
        contrib = torch.einsum("omyx,micd->yxoicd", w2, w1)        # [7,7,3,3,3,3]
        bias_tap = torch.einsum("omyx,m->yxo", w2, b1)             # [7,7,3]

        def rng(v):   # included 7x7 tap range for variant v along one axis
            if v == 0:
                return range(0, 7)
            if v <= 3:
                return range(v, 7)      # top/left taps 0..v-1 fall outside the image
            return range(0, 7 - (v - 3))
        w_eff = torch.zeros(49, 3, 3, 9, 9, dtype=torch.float64)
        b_eff = torch.zeros(49, 3, dtype=torch.float64)
        for vy in range(7):
            for vx in range(7):
                v = vy * 7 + vx
                b_eff[v] = b2 + bc
                for ky in rng(vy):
                    for kx in rng(vx):
                        w_eff[v, :, :, ky:ky + 3, kx:kx + 3] += contrib[ky, kx]
                        b_eff[v] += bias_tap[ky, kx]
        return {"prenet.w_eff": w_eff.reshape(49, 3, 3, 81).float().contiguous(),
                "prenet.b_eff": b_eff.float().contiguous(),
                "prenet.wc": wc.reshape(3, 3, 49).float().contiguous()}

    # ---- plan construction ----------------------------------------------------------------------
    def _wptr(self, key):
        e = self._packed["index"][key]
        return (self._packed["w16"].data_ptr() + 2 * e["w_off"], self._packed["b32"].data_ptr() + 4 * e["b_off"], e)

    def _tapgemm(self, key: str, src: Act, dst_ptr: int, dst_cs: int, out_h: int, out_w: int, *, stride=1,
                 relu=False, residual: Optional[Act] = None, out_mode=L.OUT_ACT16, k: Optional[int] = None,
                 m_rows: Optional[int] = None, a_cs: Optional[int] = None, bias=True) -> L.TapGemmArgs:
        wp, bp, e = self._wptr(key)
        a = L.TapGemmArgs()
        a.dtype = self._dt
        a.mode = L.TG_PATCH if stride != 1 else L.TG_FLAT
        a.ntaps = e["ntaps"]
        a.stride = stride
        a.batch, a.out_h, a.out_w = src.b, out_h, out_w
        a.in_h, a.in_w = src.h, src.w
        a.m_rows = m_rows if m_rows is not None else src.b * (out_h + 1) * (out_w + 1)
        a.k = k if k is not None else e["k"]
        a.a_cs = a_cs if a_cs is not None else src.cs
        a.n, a.n_alloc, a.w_cs = e["n"], e["n_alloc"], e["w_cs"]
        a.out_cs = dst_cs
        a.res_cs = residual.cs if residual is not None else 0
        a.relu = int(relu)
        a.out_mode = out_mode
        a.a, a.w = src.ptr, wp
        a.bias = bp if bias else None
        a.residual = residual.ptr if residual is not None else None
        a.out = dst_ptr
        return a

    @staticmethod
    def _lane_shares(nb: int, sm_count: int):
        """SM share of each branch's persistent tap-GEMMs while the branches of a stage run side by side:
        proportional to the measured marginal cost of one BasicBlock conv of that branch (tools/conv_sweep.py)."""
        w = [float(v) for v in os.environ.get("BUCTD_LANE_WEIGHTS", "40,20,10,10").split(",")][:nb]
        tot = sum(w)
        sh = [max(8, int(round(sm_count * v / tot))) for v in w]
        sh[0] += sm_count - sum(sh)
        return sh

    def _build(self, batch: int, device) -> dict:
        arch, dt = self.arch, self._dt
        sm_count = torch.cuda.get_device_properties(device).multi_processor_count if torch.cuda.is_available() else 148
        bump = _Bump()
        H4, W4 = arch.img_h // 4, arch.img_w // 4
        specs: List[tuple] = []      # (kind, args, label) with workspace-relative pointers resolved later
        acts: Dict[str, Act] = {}

        def act(name, h, w, c, cs=None) -> Act:
            cs = cs or _rup(c, 8)
            old = acts.get(name)
            if old is not None:            # same-shaped scratch is shared between modules of a stage
                assert (old.h, old.w, old.c, old.cs) == (h, w, c, cs), name
                return old
            a = Act(bump.take(batch * (h + 1) * (w + 1) * cs * 2), batch, h, w, c, cs)
            acts[name] = a
            return a

        def raw(nbytes) -> int:
            return bump.take(nbytes)

        marks: Dict[str, tuple] = {}

        def mark(name, a):
            marks[name] = (len(specs), a)

        # Lanes (include/buctd_b200.h, "Lanes"): branch i of a stage is lane i; the term chains of an exchange/fuse
        # step are spread over all lanes.  `sched` = [lane, max_ctas] of the ops appended next.
        lane_mode = os.environ.get("BUCTD_LANES", "2")          # 0: one lane; 1: all phases; 2: CoAM + fuse steps only (default:
        # BasicBlock convs of different branches slow each other down when run side by side, tools/lane_bench.py)
        use_lanes = lane_mode != "0"
        block_lanes = lane_mode == "1"
        nl_max = L.MAX_LANES if use_lanes else 1
        sched = [0, 0]

        def at(lane, share=0):
            sched[0], sched[1] = (lane % nl_max, share) if use_lanes else (0, 0)

        def emit(kind, spec, label):
            specs.append((kind, spec, label, sched[0], sched[1]))

        class _Emit:                      # list-like view for _build_coam
            @staticmethod
            def append(t):
                emit(*t)

        # BasicBlock convolutions of one branch and module (identical geometry) go out as ONE chain launch
        # (include/buctd_b200.h, buctd_tapgemm_chain) unless BUCTD_CHAIN=0 or the branches run on their own lanes.
        use_chain = os.environ.get("BUCTD_CHAIN", "1") != "0" and not block_lanes
        collect: List[Optional[list]] = [None]

        def conv(cv: Conv, src: Act, dst: Act, residual=None, relu=None, key=None):
            oh, ow = (src.h, src.w) if cv.stride == 1 else (src.h // 2, src.w // 2)
            assert (dst.h, dst.w) == (oh, ow), (cv.w, dst, oh, ow)
            spec = ("tg", key or cv.w, src, dst, dict(stride=cv.stride, relu=cv.relu if relu is None else relu, residual=residual))
            if collect[0] is not None:
                collect[0].append((spec, cv.w))
            else:
                emit(L.OP_TAPGEMM, spec, cv.w)

        # BUCTD_CHAIN_SPLIT=n cuts every chain into n crop ranges (0: as many as make the live set -- block input,
        # intermediate, output -- fit BUCTD_CHAIN_L2_MB of L2).  Crops are independent, so the result is bit-identical.
        # MEASURED (B = 64, round 2): branch 0 streams 530 MB of DRAM per launch unsplit (3 x 43.5 MB live against 126 MB),
        # but two L2-resident halves take 358 us against 262 us unsplit -- the launch is bound by its pipeline ramp and the
        # per-item dataflow, not by DRAM; whole forward 14.15 ms (auto) / 16.3 ms (2 everywhere) against 13.37.  Default 1.
        split_env = int(os.environ.get("BUCTD_CHAIN_SPLIT", "1"))
        l2_budget = int(float(os.environ.get("BUCTD_CHAIN_L2_MB", "90")) * 2 ** 20)

        def sub(a: Optional[Act], b0: int, nb_: int) -> Optional[Act]:
            if a is None:
                return None
            return Act(a.ptr + b0 * (a.h + 1) * (a.w + 1) * a.cs * 2, nb_, a.h, a.w, a.c, a.cs)

        def flush_chain(label):
            items, collect[0] = collect[0], None
            for lo in range(0, len(items), L.MAX_CHAIN):
                part = items[lo:lo + L.MAX_CHAIN]
                if len(part) == 1:
                    emit(L.OP_TAPGEMM, part[0][0], part[0][1])
                    continue
                src0 = part[0][0][2]
                live = 3 * src0.b * (src0.h + 1) * (src0.w + 1) * src0.cs * 2
                nsp = split_env if split_env > 0 else max(1, -(-live // l2_budget))
                nsp = max(1, min(nsp, src0.b))
                if nsp == 1:
                    emit(L.OP_CHAIN, ("chain", [sp for sp, _ in part], [lb for _, lb in part]), label)
                    continue
                bounds = [src0.b * i // nsp for i in range(nsp + 1)]
                for i in range(nsp):
                    b0, nb_ = bounds[i], bounds[i + 1] - bounds[i]
                    sps = []
                    for (tag, key, src, dst, kw), _ in part:
                        kw2 = dict(kw)
                        kw2["residual"] = sub(kw.get("residual"), b0, nb_)
                        sps.append((tag, key, sub(src, b0, nb_), sub(dst, b0, nb_), kw2))
                    emit(L.OP_CHAIN, ("chain", sps, [lb for _, lb in part]), label + (":crops%d" % i))

        # ---- stem ---------------------------------------------------------------------------
        s1 = act("stem1", arch.img_h // 2, arch.img_w // 2, 64)
        s2 = act("stem2", H4, W4, 64)
        conv(arch.stem[1], s1, s2)
        mark("stem", s2)
        # ---- layer1 ------------------------------------------------------------------------
        l1 = [act("l1_a", H4, W4, 256), act("l1_b", H4, W4, 256)]
        t64a, t64b = act("l1_t1", H4, W4, 64), act("l1_t2", H4, W4, 64)
        ds = act("l1_ds", H4, W4, 256)
        cur = s2
        for i, blk in enumerate(arch.layer1):
            conv(blk.convs[0], cur, t64a)
            conv(blk.convs[1], t64a, t64b)
            res = cur
            if blk.downsample is not None:
                conv(blk.downsample, cur, ds)
                res = ds
            out = l1[i % 2]
            conv(blk.convs[2], t64b, out, residual=res, relu=True)
            cur = out
        mark("layer1", cur)

        # ---- stages ------------------------------------------------------------------------
        ys = [cur]
        cond_acts = {}
        head_ops = []          # immediate ops that read the model input (stem conv1, resizes, prenet)
        for si, (trans, mods) in enumerate(zip(arch.transitions, arch.stages)):
            chans = arch.stage_channels[si]
            nb = len(chans)
            geo = [(H4 >> i, W4 >> i) for i in range(nb)]
            # three rotating buffers per branch
            bufs = [[act("s%d_b%d_%d" % (si + 2, i, j), geo[i][0], geo[i][1], chans[i]) for j in range(3)] for i in range(nb)]
            rot = [0] * nb

            def nxt(i):
                rot[i] = (rot[i] + 1) % 3
                return bufs[i][rot[i]]

            shares = self._lane_shares(nb, sm_count) if use_lanes else [0] * nb
            xs = []
            for i in range(nb):
                t = trans[i]
                at(i, shares[i])
                if t is None:
                    xs.append(ys[i])
                elif i < len(ys):
                    d = nxt(i)
                    conv(t[0], ys[i], d)
                    xs.append(d)
                else:
                    src = ys[-1]
                    for j, cv in enumerate(t):
                        last = j == len(t) - 1
                        d = nxt(i) if last else act("s%d_tr%d_%d" % (si + 2, i, j), src.h // 2, src.w // 2, cv.cout)
                        conv(cv, src, d)
                        src = d
                    xs.append(src)
            mark("transition%d" % (si + 1), list(xs))
            # make sure every branch tensor lives in this stage's rotation (needed for in-place rotation logic)
            if si in arch.coam and arch.use_attention:
                xs = self._build_coam(arch.coam[si], xs, batch, act, raw, _Emit, head_ops, nxt, cond_acts,
                                      lambda i: at(i, shares[i]))
                mark("stage%d_att" % (si + 1), list(xs))
            for mi, m in enumerate(mods):
                for i in range(nb):
                    at(i, shares[i]) if block_lanes else at(0)
                    if use_chain:
                        collect[0] = []
                    for blk in m.branches[i]:
                        tmp = nxt(i)
                        conv(blk.convs[0], xs[i], tmp)
                        out = nxt(i)
                        if out.ptr == xs[i].ptr:       # never overwrite the residual source
                            out = nxt(i)
                        conv(blk.convs[1], tmp, out, residual=xs[i], relu=True)
                        xs[i] = out
                    if use_chain:
                        flush_chain("%s.branches.%d:chain" % (m.prefix, i))
                if not m.fuse:
                    continue
                outs = []
                scratch_id = 0
                pending = []
                fuse_share = max(sm_count // 4, 1) if use_lanes else 0
                for i, terms in enumerate(m.fuse):
                    fa_terms = []
                    for t in terms:
                        at(i * nb + t.src, fuse_share)
                        if t.kind == "id":
                            fa_terms.append((xs[t.src], 0))
                        elif t.kind == "up":
                            sh = t.src - i
                            d = act("s%d_up%d_%d" % (si + 2, i, t.src), geo[t.src][0], geo[t.src][1], chans[i])
                            conv(t.convs[0], xs[t.src], d)
                            fa_terms.append((d, sh))
                        else:
                            src = xs[t.src]
                            for j, cv in enumerate(t.convs):
                                d = act("s%d_dn%d_%d_%d" % (si + 2, i, t.src, j), src.h // 2, src.w // 2, cv.cout)
                                conv(cv, src, d)
                                src = d
                            fa_terms.append((src, 0))
                    pending.append((i, fa_terms))
                # all fuse outputs read the pre-fuse xs: write them to fresh rotation slots
                new_xs = []
                for i, fa_terms in pending:
                    out = nxt(i)
                    if out.ptr == xs[i].ptr:
                        out = nxt(i)
                    at(i) if block_lanes else at(0)
                    emit(L.OP_FUSE, ("fuse", fa_terms, out), "%s.fuse%d" % (m.prefix, i))
                    new_xs.append(out)
                xs = new_xs
            mark("stage%d" % (si + 2), list(xs))
            ys = xs
        # ATT_MODULES[3]: stage4_att on y_list[0] after stage4 (pose_hrnet_coam.py:560-562); nxt / shares are stage 4's
        if 3 in arch.coam and arch.use_attention:
            ys = self._build_coam(arch.coam[3], ys[:1], batch, act, raw, _Emit, head_ops, nxt, cond_acts,
                                  lambda i: at(i, shares[i])) + list(ys[1:])
            mark("stage4_att", list(ys[:1]))
        at(0)
        final_src = ys[0]
        nbytes = bump.off
        return {"specs": specs, "acts": acts, "marks": marks, "bytes": nbytes, "final_src": final_src, "s1": s1,
                "head_ops": head_ops, "cond_acts": cond_acts}

    def _build_coam(self, das: List[DAModule], xs, batch, act, raw, specs, head_ops, nxt, cond_acts, at):
        arch = self.arch
        outs = []
        for i, (da, x) in enumerate(zip(das, xs)):
            at(i)
            C_, cc, h, w = da.c, da.cc, da.h, da.w
            assert (x.h, x.w, x.c) == (h, w, C_)
            ccs = _rup(cc, 16)
            cond = act("cond%d" % i, h, w, cc, ccs)
            cond_acts[i] = cond
            head_ops.append(("resize", cond, cc))
            p = da.prefix
            # channel attention
            fuse_cnn = not arch.att_channel_only and os.environ.get("BUCTD_COAM_CNN2", "1") != "0"
            if fuse_cnn:
                # channel_attention_module.cnn and position_attention_module.cnn read the same x: ONE 3x3 tap-GEMM with N = 2C;
                # y / yp are the two channel halves of its output (row stride 2C)
                yy = act(p + ":y|yp", h, w, 2 * C_)
                y = Act(yy.ptr, yy.b, h, w, C_, yy.cs)
                specs.append((L.OP_TAPGEMM, ("tg", p + ":cnn2", x, yy, dict(stride=1, relu=False, residual=None)), p + ":cnn2"))
            else:
                y = act(p + ":y", h, w, C_)
                specs.append((L.OP_TAPGEMM, ("tg", da.chan_cnn.w, x, y, dict(stride=1, relu=False, residual=None)), da.chan_cnn.w))
            yc = act(p + ":yc", h, w, C_)
            specs.append((L.OP_TAPGEMM, ("tg", da.chan_cnn_cond.w, cond, yc, dict(stride=1, relu=False, residual=None, k=cc)),
                          da.chan_cnn_cond.w))
            hws = _rup(h * w, 8)          # planar row stride: 16-byte multiples for TMA (H*W = 108 at the 1/32 scale of 384x288)
            rows1, ldp = (h + 1) * (w + 1), _rup((h + 1) * (w + 1), 8)     # channel-attention output in padded pixel order
            planar = raw(batch * C_ * ldp * 2)
            scratch = raw(L.chan_att_scratch_floats(batch, h, w, C_) * 4)
            c_out = raw(batch * C_ * hws * 2)
            specs.append((L.OP_CHAN_ATT, ("chan", y, yc, scratch, planar, ldp), p + ":chan_att"))
            specs.append((L.OP_TAPGEMM, ("lin_planar", da.chan_fc_o.w, planar, c_out, batch * C_, rows1, ldp, hws), da.chan_fc_o.w))
            p_out = None
            if not arch.att_channel_only:
                yp = Act(yy.ptr + 2 * C_, yy.b, h, w, C_, yy.cs) if fuse_cnn else act(p + ":yp", h, w, C_)
                ycp = act(p + ":ycp", h, w, cc, ccs)
                q = act(p + ":q", h, w, C_)
                kv = act(p + ":kv", h, w, 2 * C_)
                o = act(p + ":o", h, w, C_)
                p_out = act(p + ":p_out", h, w, C_)
                if not fuse_cnn:
                    specs.append((L.OP_TAPGEMM, ("tg", da.pos_cnn.w, x, yp, dict(stride=1, relu=False, residual=None)), da.pos_cnn.w))
                specs.append((L.OP_TAPGEMM, ("tg", da.pos_cnn_cond.w, cond, ycp, dict(stride=1, relu=False, residual=None, k=cc)),
                              da.pos_cnn_cond.w))
                specs.append((L.OP_TAPGEMM, ("tg", da.fc_q.w, ycp, q, dict(stride=1, relu=False, residual=None, k=cc)), da.fc_q.w))
                specs.append((L.OP_TAPGEMM, ("tg", p + ":kv", yp, kv, dict(stride=1, relu=False, residual=None)), p + ":kv"))
                specs.append((L.OP_POS_ATT, ("pos", q, kv, o, C_), p + ":pos_att"))
                specs.append((L.OP_TAPGEMM, ("tg", da.fc_o.w, o, p_out, dict(stride=1, relu=False, residual=None)), da.fc_o.w))
            out = nxt(i)
            if out.ptr == x.ptr:
                out = nxt(i)
            specs.append((L.OP_DA_COMBINE, ("da", x, p_out, c_out, out, hws), p + ":combine"))
            outs.append(out)
        return outs

    def _materialise(self, batch: int, device) -> dict:
        if self._packed is None or self._packed["device"] != device:
            self._packed = self._pack(device)
        bld = self._build(batch, device)
        ws = torch.zeros(bld["bytes"] + 1024, dtype=torch.uint8, device=device)
        base = _rup(ws.data_ptr(), 1024)
        ctx = L.Context.get(device.index)
        plan = L.Plan(ctx)

        def fix(a: Act) -> Act:
            return Act(base + a.ptr, a.b, a.h, a.w, a.c, a.cs)

        for kind, spec, label, lane, share in bld["specs"]:
            tag = spec[0]
            if tag == "tg":
                _, key, src, dst, kw = spec
                kw = dict(kw)
                res = kw.pop("residual")
                oh, ow = dst.h, dst.w
                args = self._tapgemm(key, fix(src), base + dst.ptr, dst.cs, oh, ow,
                                     residual=fix(res) if res is not None else None, **kw)
            elif tag == "chain":
                _, tgs, _labels = spec
                args = L.TapGemmChainArgs()
                args.n_layers = len(tgs)
                for li, (_, key, src, dst, kw) in enumerate(tgs):
                    kw = dict(kw)
                    res = kw.pop("residual")
                    args.layer[li] = self._tapgemm(key, fix(src), base + dst.ptr, dst.cs, dst.h, dst.w,
                                                   residual=fix(res) if res is not None else None, **kw)
            elif tag == "lin_planar":
                _, key, src_off, dst_off, rows, k_in, ld_in, hws = spec
                a = Act(base + src_off, 1, 0, 0, k_in, ld_in)
                args = self._tapgemm(key, a, base + dst_off, hws, 0, 0, out_mode=L.OUT_ROWS16, m_rows=rows, a_cs=ld_in)
                args.batch = 1
            elif tag == "fuse":
                _, terms, out = spec
                args = L.FuseArgs()
                args.dtype, args.batch, args.h, args.w, args.c, args.out_cs = self._dt, batch, out.h, out.w, out.c, out.cs
                args.nterms, args.relu = len(terms), 1
                for j, (t, sh) in enumerate(terms):
                    args.term[j] = base + t.ptr
                    args.term_cs[j] = t.cs
                    args.term_shift[j] = sh
                args.out = base + out.ptr
            elif tag == "chan":
                _, y, yc, scratch, planar, hws = spec
                args = L.ChanAttArgs()
                args.dtype, args.batch, args.h, args.w, args.c = self._dt, batch, y.h, y.w, y.c
                args.y_cs, args.yc_cs, args.out_stride, args.out_padded = y.cs, yc.cs, hws, 1
                args.y, args.yc, args.scratch, args.out = base + y.ptr, base + yc.ptr, base + scratch, base + planar
            elif tag == "pos":
                _, q, kv, o, c = spec
                args = L.PosAttArgs()
                args.dtype, args.batch, args.h, args.w, args.c = self._dt, batch, q.h, q.w, c
                args.q_cs, args.k_cs, args.v_cs, args.o_cs = q.cs, kv.cs, kv.cs, o.cs
                args.q, args.k, args.v, args.o = base + q.ptr, base + kv.ptr, base + kv.ptr + 2 * c, base + o.ptr
            elif tag == "da":
                _, x, p_out, c_out, out, hws = spec
                args = L.DaCombineArgs()
                args.c_stride = hws
                args.dtype, args.batch, args.h, args.w, args.c = self._dt, batch, x.h, x.w, x.c
                args.x_cs, args.p_cs, args.out_cs = x.cs, (p_out.cs if p_out is not None else 0), out.cs
                args.channel_only = int(self.arch.att_channel_only)
                args.x, args.c_out, args.out = base + x.ptr, base + c_out, base + out.ptr
                args.p_out = base + p_out.ptr if p_out is not None else None
            else:
                raise AssertionError(tag)
            plan.add(kind, args, label, lane=lane, max_ctas=share)
        if os.environ.get("BUCTD_GRAPH", "1") != "0":
            plan.instantiate()
        marks = {}
        for name, (idx, a) in bld["marks"].items():
            marks[name] = (idx, [fix(t) for t in a] if isinstance(a, list) else fix(a))
        return {"plan": plan, "ws": ws, "base": base, "ctx": ctx, "s1": fix(bld["s1"]),
                "final_src": fix(bld["final_src"]), "marks": marks,
                "cond": {i: fix(a) for i, a in bld["cond_acts"].items()}, "head_ops": bld["head_ops"],
                "acts": {k: fix(v) for k, v in bld["acts"].items()}}

    # ---- forward --------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, _stop_after: Optional[str] = None) -> torch.Tensor:
        """The reference's ``forward(x[B, 3+Cc, H, W] float32)`` (pose_hrnet_coam.py:488-572, pose_hrnet.py:445-495):
        RGB planes + condition planes in one tensor, any device."""
        if self.training:
            raise NotImplementedError("model.train(): use buctd_b200.train.TrainStep (the training step has its own plan); "
                                      "forward() is the inference path (model.eval())")
        if not torch.cuda.is_available():
            raise L.BuctdError("no CUDA device: buctd_b200 has no CPU path")
        arch = self.arch
        x = x.cuda()                                                  # pose_hrnet_coam.py:495
        if (arch.use_attention or arch.use_prenet) and x[:, 3:].shape[1] == 0:
            raise Exception("condition is empty, please check your dataloader")    # :498-499
        x = x.contiguous().float()
        B, Cx, H, W = x.shape
        if (H, W) != (arch.img_h, arch.img_w):
            raise ValueError("input is %dx%d but MODEL.IMAGE_SIZE is %dx%d (CoAM's HWxHW projection fixes the size)"
                             % (W, H, arch.img_w, arch.img_h))
        need = 3 + arch.in_cond
        if Cx < need or (arch.in_cond == 0 and Cx != 3):
            raise ValueError("expected %d input channels, got %d" % (need, Cx))
        return self._run(x, Cx, None, _stop_after)

    def forward_from_proposals(self, rgb: torch.Tensor, kpts: torch.Tensor, colors=None) -> torch.Tensor:
        """Fast-path entry for the conditional top-down call: the normalised crop ``rgb`` [B, 3, H, W] (float32 or
        float16, host -- ideally pinned -- or device) and the bottom-up proposals ``kpts`` [B, K, >=2] (float32 crop
        pixels, the ``cond_joints`` the dataset rasterises, lib/dataset/JointsDataset.py:314-326).  The condition is
        rasterised on the device: 204 B of keypoints per crop cross PCIe instead of 17 float32 planes (7.5 MB).  For the
        stacked condition of CoAM the three CoAM scales are written directly from the keypoints
        (buctd_cond_from_kpts); the other modes rasterise at full resolution on the device and take ``forward``.
        Same result as ``forward(cat(rgb, rasterised condition))`` (heat-maps [B, K, H/4, W/4] float32 on CUDA)."""
        if self.training:
            raise NotImplementedError("forward_from_proposals is an inference entry point (model.eval())")
        if not torch.cuda.is_available():
            raise L.BuctdError("no CUDA device: buctd_b200 has no CPU path")
        arch = self.arch
        if not (arch.use_attention or arch.use_prenet):
            raise ValueError("forward_from_proposals needs a conditional model (CoAM or preNet)")
        if rgb.dim() != 4 or rgb.shape[1] != 3 or tuple(rgb.shape[2:]) != (arch.img_h, arch.img_w):
            raise ValueError("rgb must be [B, 3, %d, %d], got %s" % (arch.img_h, arch.img_w, tuple(rgb.shape)))
        if rgb.dtype not in (torch.float32, torch.float16):
            raise ValueError("rgb must be float32 or float16")
        B = rgb.shape[0]
        if kpts.dim() != 3 or kpts.shape[0] != B or kpts.shape[2] < 2:
            raise ValueError("kpts must be [B, K, >=2]")
        rgb = rgb.cuda(non_blocking=True).contiguous()
        kp = kpts.cuda(non_blocking=True).float().contiguous()
        mode = "stacked" if arch.cond_stacked else ("colored" if arch.cond_colored or arch.use_prenet else "plain")
        H, W = arch.img_h, arch.img_w
        direct = (arch.use_attention and mode == "stacked" and self.antialias and H % 16 == 0 and W % 16 == 0
                  and kp.shape[1] <= 32 and sorted(arch.coam) and max(len(v) for v in arch.coam.values()) <= 3
                  and os.environ.get("BUCTD_COND_DIRECT", "1") != "0")
        if direct:
            return self._run(rgb, 3, kp, None)
        from . import condition
        x = torch.empty(B, 3 + arch.in_cond, H, W, dtype=torch.float32, device=rgb.device)
        x[:, :3].copy_(rgb)
        condition.rasterize_condition(kp, H, W, mode, colors=colors, out=x, c_off=3)
        return self._run(x, x.shape[1], None, None)

    def predict_from_proposals(self, rgb: torch.Tensor, kpts: torch.Tensor, center=None, scale=None, post_process: bool = True):
        """``forward_from_proposals`` + ``get_final_preds`` with the head fused into the decode (buctd_head_decode): the
        [B, K, H/4, W/4] heat-maps are never written.  Returns (preds [B, K, 2], maxvals [B, K, 1]) CUDA tensors -- in image
        coordinates when ``center`` / ``scale`` [B, 2] are given, else heat-map coordinates.  (stacked-condition CoAM, the
        direct condition pyramid; other modes: forward_from_proposals + inference.decode_device.)"""
        if self.training:
            raise NotImplementedError("predict_from_proposals is an inference entry point (model.eval())")
        arch = self.arch
        rgb_d = rgb.cuda(non_blocking=True).contiguous()
        kp = kpts.cuda(non_blocking=True).float().contiguous()
        H, W = arch.img_h, arch.img_w
        direct = (arch.use_attention and arch.cond_stacked and self.antialias and H % 16 == 0 and W % 16 == 0 and kp.shape[1] <= 32
                  and max(len(v) for v in arch.coam.values()) <= 3 and arch.num_joints <= 32 and arch.final.k == 1
                  and arch.final.cin in (32, 48, 64))
        if not direct or rgb_d.dtype not in (torch.float32, torch.float16) or tuple(rgb_d.shape[1:]) != (3, H, W):
            from . import inference
            return inference.decode_device(self.forward_from_proposals(rgb, kpts), center, scale, post_process=post_process)
        c = torch.as_tensor(center, dtype=torch.float32).to(rgb_d.device).contiguous() if center is not None else None
        s = torch.as_tensor(scale, dtype=torch.float32).to(rgb_d.device).contiguous() if scale is not None else None
        return self._run(rgb_d, 3, kp, None, decode=(c, s, bool(post_process)))

    def _run(self, x: torch.Tensor, Cx: int, kp: Optional[torch.Tensor], _stop_after: Optional[str], decode=None) -> torch.Tensor:
        """x: device tensor whose planes [0, 3) are the RGB crop (float32, or float16 with kp given) and, when ``kp`` is
        None, planes [3, 3 + Cc) the condition; kp: proposals for the direct condition pyramid."""
        arch = self.arch
        B, _, H, W = x.shape
        device = x.device
        key = (B, device.index)
        with torch.cuda.device(device):
            st = self._plans.get(key)
            if st is None:
                st = self._materialise(B, device)
                self._plans[key] = st
            stream = torch.cuda.current_stream(device).cuda_stream
            ctx: L.Context = st["ctx"]
            f32 = self._packed["f32"]
            stem_x, stem_c = x, Cx
            if arch.use_prenet:
                pre = torch.empty(B, 3, H, W, dtype=torch.float32, device=device)
                pa = L.PrenetArgs()
                pa.batch, pa.x_channels, pa.h, pa.w = B, Cx, H, W
                pa.x, pa.w_eff, pa.b_eff, pa.wc, pa.out = (x.data_ptr(), f32["prenet.w_eff"].data_ptr(),
                                                            f32["prenet.b_eff"].data_ptr(), f32["prenet.wc"].data_ptr(),
                                                            pre.data_ptr())
                ctx.run(L.OP_PRENET, pa, stream)
                stem_x, stem_c = pre, 3
            sa = L.StemArgs()
            s1: Act = st["s1"]
            sa.dtype, sa.batch, sa.x_channels, sa.h, sa.w, sa.out_cs = self._dt, B, stem_c, H, W, s1.cs
            sa.x_half = int(stem_x.dtype == torch.float16)
            sa.x, sa.weight, sa.bias, sa.out = stem_x.data_ptr(), f32["conv1.w"].data_ptr(), f32["conv1.b"].data_ptr(), s1.ptr
            ctx.run(L.OP_STEM, sa, stream)
            conds = st["cond"] if arch.use_attention else {}
            three = (len(conds) == 3 and self.antialias and H % 16 == 0 and W % 16 == 0
                     and all((conds[i].h, conds[i].w) == (H >> (2 + i), W >> (2 + i)) for i in range(3)))
            if kp is not None:
                # the CoAM scales straight from the proposals; a placement with fewer than three branches gets scratch maps
                ck = L.CondKptsArgs()
                ck.dtype, ck.batch, ck.k, ck.kstride, ck.h, ck.w = self._dt, B, kp.shape[1], kp.shape[2], H, W
                spare = []
                for i in range(3):
                    if i in conds:
                        ck.out_cs[i], ck.out[i] = conds[i].cs, conds[i].ptr
                    else:
                        cs = _rup(kp.shape[1], 16)
                        t = torch.empty(B * ((H >> (2 + i)) + 1) * ((W >> (2 + i)) + 1) * cs, dtype=self._torch_dt, device=device)
                        spare.append(t)
                        ck.out_cs[i], ck.out[i] = cs, t.data_ptr()
                ck.kpts = kp.data_ptr()
                ctx.call("buctd_cond_from_kpts", ck, stream)
                self._keepalive_kp = (kp, spare)
            elif three and W <= 288 and os.environ.get("BUCTD_RESIZE3", "1") != "0":
                r3 = L.Resize3Args()               # the three CoAM scales in one pass over the condition planes
                r3.dtype, r3.batch, r3.x_channels, r3.c_off, r3.cc, r3.h, r3.w = self._dt, B, Cx, 3, conds[0].c, H, W
                for i in range(3):
                    r3.out_cs[i] = conds[i].cs
                    r3.out[i] = conds[i].ptr
                r3.x = x.data_ptr()
                ctx.call("buctd_cond_resize3", r3, stream)
            elif arch.use_attention:
                for i, cond in st["cond"].items():
                    ra = L.ResizeArgs()
                    ra.dtype, ra.batch, ra.x_channels, ra.c_off, ra.cc = self._dt, B, Cx, 3, cond.c
                    ra.h, ra.w, ra.oh, ra.ow, ra.out_cs, ra.antialias = H, W, cond.h, cond.w, cond.cs, int(self.antialias)
                    ra.x, ra.out = x.data_ptr(), cond.ptr
                    ctx.run(L.OP_RESIZE, ra, stream)
            last = None
            if _stop_after is not None:
                last = st["marks"][_stop_after][0]
            st["plan"].run(stream, 0, last)
            if _stop_after is not None:
                return self._export(st, st["marks"][_stop_after][1], stream)
            fs: Act = st["final_src"]
            if decode is not None:                         # head fused into the decode: no heat-maps
                center, scale, pp = decode
                K = arch.num_joints
                preds = torch.empty(B, K, 2, dtype=torch.float32, device=device)
                maxvals = torch.empty(B, K, 1, dtype=torch.float32, device=device)
                wp, bp, e = self._wptr("final_layer")
                hd = L.HeadDecodeArgs()
                hd.dtype, hd.batch, hd.k, hd.h, hd.w, hd.c, hd.a_cs, hd.w_cs, hd.post_process = self._dt, B, K, fs.h, fs.w, fs.c, fs.cs, e["w_cs"], int(pp)
                hd.feat, hd.weight, hd.bias = fs.ptr, wp, bp
                hd.center, hd.scale = (center.data_ptr() if center is not None else None), (scale.data_ptr() if scale is not None else None)
                hd.preds, hd.maxvals = preds.data_ptr(), maxvals.data_ptr()
                scratch = torch.empty(B * ((fs.h * fs.w + 255) // 256) * K * 2, dtype=torch.float32, device=device)
                hd.scratch = scratch.data_ptr()
                ctx.call("buctd_head_decode", hd, stream)
                self._keepalive = (x, stem_x, center, scale, scratch)
                return preds, maxvals
            out = torch.empty(B, arch.num_joints, fs.h, fs.w, dtype=torch.float32, device=device)
            fa = self._tapgemm("final_layer", fs, out.data_ptr(), 0, fs.h, fs.w, out_mode=L.OUT_NCHW32)
            ctx.run(L.OP_TAPGEMM, fa, stream)
            self._keepalive = (x, stem_x)      # inputs must outlive the enqueued kernels on this stream
        return out

    def _export(self, st, a, stream):
        """act16 -> nchw32 torch tensors (stage-boundary parity tests)."""
        if isinstance(a, list):
            return [self._export(st, t, stream) for t in a]
        out = torch.empty(a.b, a.c, a.h, a.w, dtype=torch.float32, device=st["ws"].device)
        st["ctx"].act16_to_nchw32(self._dt, a.ptr, a.b, a.h, a.w, a.c, a.cs, out.data_ptr(), stream)
        return out

    def launches_per_forward(self, batch: int, device_index: int = 0) -> int:
        st = self._plans.get((batch, device_index))
        if st is None:
            return 0
        n = len(st["plan"]) + 2 + len(st["cond"]) + (1 if self.arch.use_prenet else 0)
        n += 2 * sum(1 for l in st["plan"].labels if l.endswith(":chan_att"))   # scores + softmax + apply
        return n
