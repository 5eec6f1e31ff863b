"""TEST INFRASTRUCTURE -- CPU fp32 restatement of the reference forward.  Not a product path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.

A plain functional torch-fp32 restatement of ``PoseHighResolutionNet.forward`` for the two
BUCTD variants, driven directly by a reference ``state_dict`` (no nn.Module tree, NCHW fp32,
eval-mode BatchNorm).  Each function cites the reference lines it follows.  It is pinned against
the reference's own modules run on the CPU: oracle/make_golden.py writes tests/golden/*.npz from
the reference and tests/test_oracle_vs_golden.py checks this file against them (and against the
live reference when /root/reference is present).

``act`` is an optional activation hook (identity for the oracle).  The design experiments in
DESIGN.md use it to emulate bf16 storage of activations; it is not used by any test oracle.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
_ID = lambda t: t  # noqa: E731


_AQ = [_ID]   # design-experiment hook: quantiser applied to conv / linear INPUTS (identity for the oracle)

# Training-mode switches (oracle/train_oracle.py sets them around a forward): nn.BatchNorm2d with batch statistics
# (momentum BN_MOMENTUM = 0.1, pose_hrnet_coam.py:20) and nn.Dropout(0.1) on both attention maps
# (lib/models/self_attention.py:38,84,115,156) with an externally supplied keep-mask so that the CUDA path can be checked.
_CTX = {"train": False, "running": None, "drop": None, "drop_p": 0.0}


def _conv(sd: SD, p: str, x, stride=1, wq=_ID):
    x = _AQ[0](x)
    # nn.Conv2d(k, stride, padding=k//2)  (pose_hrnet_coam.py:24-28, 68-74; padding='same' in the
    # preNet, pose_hrnet.py:431-442, is the same thing for odd k and stride 1)
    w = sd[p + ".weight"]
    return F.conv2d(x, wq(w), sd.get(p + ".bias"), stride=stride, padding=w.shape[-1] // 2)


def _bn(sd: SD, p: str, x, eps=1e-5):
    if _CTX["train"]:
        rm, rv = sd[p + ".running_mean"].detach().clone(), sd[p + ".running_var"].detach().clone()
        y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training=True, momentum=0.1, eps=eps)
        if _CTX["running"] is not None:
            _CTX["running"][p] = (rm, rv)
        return y
    # eval-mode nn.BatchNorm2d: (x - running_mean) / sqrt(running_var + eps) * weight + bias
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=eps)


def _basic_block(sd, p, x, act, wq):
    # BasicBlock.forward, pose_hrnet_coam.py:44-60
    out = act(F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, wq=wq))))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, wq=wq))
    return act(F.relu(out + x))


def _bottleneck(sd, p, x, act, wq):
    # Bottleneck.forward, pose_hrnet_coam.py:81-101
    out = act(F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, wq=wq))))
    out = act(F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, wq=wq))))
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out, wq=wq))
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, wq=wq))
    return act(F.relu(out + res))


def _hr_module(sd, p, xs: List[torch.Tensor], nblocks: int, multi_scale: bool, act, wq):
    # HighResolutionModule.forward, pose_hrnet_coam.py:250-268; fuse layers :190-245
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        for k in range(nblocks):
            xs[b] = _basic_block(sd, "%s.branches.%d.%d" % (p, b, k), xs[b], act, wq)
    if nb == 1:
        return xs
    outs = []
    for i in range(nb if multi_scale else 1):
        y = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = _bn(sd, fp + ".1", _conv(sd, fp + ".0", xs[j], wq=wq))
                t = F.interpolate(act(t), scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _bn(sd, "%s.%d.1" % (fp, k), _conv(sd, "%s.%d.0" % (fp, k), t, stride=2, wq=wq))
                    t = act(F.relu(t)) if k != i - j - 1 else act(t)
            y = t if y is None else y + t
        outs.append(act(F.relu(y)))
    return outs


def _transition(sd, name, pre: List[torch.Tensor], nb_cur: int, act, wq):
    # _make_transition_layer :395-434 and its use in forward :513-518, 527-532, 541-546:
    # existing branches pass through (or 3x3 s1 conv if the channel count changes); every new
    # branch is a chain of 3x3 s2 conv+BN+ReLU applied to the LAST previous branch.
    out = []
    for i in range(nb_cur):
        if i < len(pre):
            p = "%s.%d" % (name, i)
            if (p + ".0.weight") in sd:
                out.append(act(F.relu(_bn(sd, p + ".1", _conv(sd, p + ".0", pre[i], wq=wq)))))
            else:
                out.append(pre[i])
        else:
            t = pre[-1]
            for j in range(i + 1 - len(pre)):
                p = "%s.%d.%d" % (name, i, j)
                t = act(F.relu(_bn(sd, p + ".1", _conv(sd, p + ".0", t, stride=2, wq=wq))))
            out.append(t)
    return out


def resize_aa(cond: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """``TF.resize(cond, (h, w))`` as torchvision 0.26 does it for tensors: bilinear,
    align_corners=False, antialias=True (pose_hrnet_coam.py:755).  Written
    out as the separable triangle filter so the oracle does not depend on torchvision."""
    def weights(n_in, n_out):
        scale = n_in / n_out
        support = scale if scale >= 1.0 else 1.0
        m = torch.zeros(n_out, n_in, dtype=torch.float64)
        for i in range(n_out):
            center = scale * (i + 0.5)
            lo = max(0, int(center - support + 0.5))
            hi = min(n_in, int(center + support + 0.5))
            js = torch.arange(lo, hi, dtype=torch.float64)
            wt = (1.0 - ((js - center + 0.5) / (scale if scale >= 1.0 else 1.0)).abs()).clamp(min=0)
            m[i, lo:hi] = wt / wt.sum()
        return m
    wh = weights(cond.shape[2], h).to(device=cond.device, dtype=cond.dtype)
    ww = weights(cond.shape[3], w).to(device=cond.device, dtype=cond.dtype)
    # horizontal pass first, then vertical (ATen's separable order); fp32 like the reference
    t = torch.einsum("bchw,xw->bchx", cond, ww)
    return torch.einsum("bchx,yh->bcyx", t, wh)


def _lin(sd, p, x, wq=_ID):
    x = _AQ[0](x)
    return F.linear(x, wq(sd[p + ".weight"]), sd[p + ".bias"])


def _dropout(att, tag):
    """nn.Dropout on an attention map in training mode: keep-mask from the hook, kept values / (1 - p)"""
    if not _CTX["train"] or _CTX["drop_p"] <= 0.0 or _CTX["drop"] is None:
        return att
    keep = _CTX["drop"](tag, att)
    return att * keep / (1.0 - _CTX["drop_p"])


def _da_module(sd, p, x, cond, channel_only: bool, act, wq, rec=None):
    # DAModule.forward :705-724
    bs, c, h, w = x.shape
    cp = p + ".channel_attention_module"
    # ChannelAttentionModule.forward :673-689 + SimplifiedScaledDotProductAttention :133-160 (h=1)
    y = act(_conv(sd, cp + ".cnn", x, wq=wq)).view(bs, c, -1)                   # [B,C,HW] keys = values
    yc = act(_conv(sd, cp + ".cnn_cond", cond, wq=wq)).view(bs, c, -1)          # [B,C,HW] queries
    att = torch.softmax(torch.matmul(yc, y.transpose(1, 2)) / math.sqrt(h * w), -1)  # [B,C,C]
    c_out = act(_lin(sd, cp + ".pa.fc_o", act(torch.matmul(_dropout(att, p + ":chan"), y)), wq)).view(bs, c, h, w)
    if rec is not None:
        rec(p + ":chan_att", att)
        rec(p + ":c_out", c_out)
    if channel_only:
        return act(x * c_out)
    pp = p + ".position_attention_module"
    # PositionAttentionModule.forward :645-660 + ScaledDotProductAttention :61-88 (h=1, eval: no dropout)
    yp = act(_conv(sd, pp + ".cnn", x, wq=wq)).view(bs, c, -1).permute(0, 2, 1)           # [B,HW,C]
    ycp = act(_conv(sd, pp + ".cnn_cond", cond, wq=wq))
    ycp = ycp.view(bs, ycp.shape[1], -1).permute(0, 2, 1)                                 # [B,HW,Cc]
    q = act(_lin(sd, pp + ".pa.fc_q", ycp, wq))
    k = act(_lin(sd, pp + ".pa.fc_k", yp, wq))
    v = act(_lin(sd, pp + ".pa.fc_v", yp, wq))
    att = torch.softmax(torch.matmul(q, k.transpose(1, 2)) / math.sqrt(c), -1)            # [B,HW,HW]
    o = act(_lin(sd, pp + ".pa.fc_o", act(torch.matmul(act(_dropout(att, p + ":pos")), v)), wq))                # [B,HW,C]
    p_out = o.permute(0, 2, 1).reshape(bs, c, h, w)
    if rec is not None:
        rec(p + ":pos_att_rows", att[:, ::61])
        rec(p + ":q", q)
        rec(p + ":k", k)
        rec(p + ":p_out", p_out)
    return act(x + (p_out + c_out))


def _coam_block(sd, name, xs, cond_hm, stacked: bool, colored: bool, channel_only: bool, act, wq,
                resize: Callable = resize_aa, rec=None):
    # CoAMBlock.forward :750-757
    resize = _CTX.get("resize") or resize        # tests: TF.resize without antialiasing (torchvision 0.9 behaviour)
    if not colored and not stacked:
        cond_hm = cond_hm[:, 0:1]
    out = []
    for i, x in enumerate(xs):
        cond_i = act(resize(cond_hm, x.shape[2], x.shape[3]))
        out.append(_da_module(sd, "%s.att_layers.%d" % (name, i), x, cond_i, channel_only, act, wq, rec))
    return out


def forward(sd: SD, cfg, x: torch.Tensor, act: Callable = _ID, wq: Callable = _ID,
            taps: Optional[dict] = None) -> torch.Tensor:
    """x [B, 3+Cc, H, W] fp32 -> heatmaps [B, K, H/4, W/4] fp32.
    pose_hrnet_coam.py:488-572 (CoAM) and pose_hrnet.py:445-495 (plain / preNet).
    ``taps`` (optional dict) receives intermediate tensors at stage boundaries."""
    model, extra = cfg["MODEL"], cfg["MODEL"]["EXTRA"]
    name = model["NAME"]
    use_att = name == "pose_hrnet_coam" and bool(extra.get("USE_ATTENTION", False))
    use_pre = name == "pose_hrnet" and bool(extra.get("USE_PRE_NET", False))
    att_cfg = list(model.get("ATT_MODULES", [False] * 4)) if use_att else [False] * 4
    stacked = bool(cfg["DATASET"].get("STACKED_CONDITION", False))
    colored = bool(cfg["DATASET"].get("COLORED", False))
    ch_only = bool(model.get("ATT_CHANNEL_ONLY", False))
    rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)

    x = x.float()
    cond_hm = None
    if use_att:
        if x[:, 3:].shape[1] == 0:
            raise Exception("condition is empty, please check your dataloader")   # :498-499
        x, cond_hm = x[:, :3], x[:, 3:]
    elif use_pre:
        if x[:, 3:].shape[1] == 0:
            raise Exception("condition is empty, please check your dataloader")   # pose_hrnet.py:454-455
        # rgb: conv3x3(3->64)+BN, conv7x7(64->3)+BN ; cond: conv7x7(3->3)+BN ; sum, no ReLU (:452-458)
        x0 = act(_bn(sd, "rgb_preNet.1", _conv(sd, "rgb_preNet.0", x[:, :3], wq=wq)))
        x0 = _bn(sd, "rgb_preNet.3", _conv(sd, "rgb_preNet.2", x0, wq=wq))
        x1 = _bn(sd, "cond_preNet.1", _conv(sd, "cond_preNet.0", x[:, 3:], wq=wq))
        x = act(x0 + x1)
        rec("prenet", x)

    x = act(F.relu(_bn(sd, "bn1", _conv(sd, "conv1", x, stride=2, wq=wq))))       # :505-507
    x = act(F.relu(_bn(sd, "bn2", _conv(sd, "conv2", x, stride=2, wq=wq))))       # :508-510
    rec("stem", x)
    for i in range(4):                                                            # :511
        x = _bottleneck(sd, "layer1.%d" % i, x, act, wq)
    rec("layer1", x)

    ys = [x]
    for sidx, att_slot in ((2, 0), (3, 1), (4, 2)):
        scfg = extra["STAGE%d" % sidx]
        nb, nmod, nblk = int(scfg["NUM_BRANCHES"]), int(scfg["NUM_MODULES"]), int(scfg["NUM_BLOCKS"][0])
        xs = _transition(sd, "transition%d" % (sidx - 1), ys, nb, act, wq)
        rec("transition%d" % (sidx - 1), xs)
        if att_cfg[att_slot]:
            xs = _coam_block(sd, "stage%d_att" % (att_slot + 1), xs, cond_hm, stacked, colored, ch_only, act, wq,
                             rec=rec if taps is not None else None)
            rec("stage%d_att" % (att_slot + 1), xs)
        for m in range(nmod):
            multi = not (sidx == 4 and m == nmod - 1)
            xs = _hr_module(sd, "stage%d.%d" % (sidx, m), xs, nblk, multi, act, wq)
        rec("stage%d" % sidx, xs)
        ys = xs
    if att_cfg[3]:
        ys = _coam_block(sd, "stage4_att", ys[:1], cond_hm, stacked, colored, ch_only, act, wq)
    out = _conv(sd, "final_layer", ys[0], wq=wq)                                  # :565
    return out
