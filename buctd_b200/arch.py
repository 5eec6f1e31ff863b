"""Static description of the BUCTD top-down network (HRNet backbone + CoAM / preNet + head).

This is the single place that knows the layer graph and the reference's ``state_dict`` key
names.  It is pure data: the parameter container (``module.py``), the weight packer and the
plan builder (``plan.py``) all walk it.  Structure and key names follow the reference's
module tree so that ``load_state_dict(strict=True)`` of a reference checkpoint works:

* stem / layer1 / transitions / stages / head: /root/reference/lib/models/pose_hrnet_coam.py:277-334
  (``_make_layer`` :436-453, ``_make_transition_layer`` :395-434, ``_make_stage`` :455-485,
  ``HighResolutionModule._make_fuse_layers`` :190-245)
* CoAM: ``CoAMBlock`` :728-757, ``DAModule`` :692-725, ``PositionAttentionModule`` :631-660,
  ``ChannelAttentionModule`` :662-689, attention projections
  /root/reference/lib/models/self_attention.py:16-38,100-115
* preNet: /root/reference/lib/models/pose_hrnet.py:282-284,431-442
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Tuple

BN_EPS = 1e-5


@dataclass
class Conv:
    w: str                 # state_dict prefix of the nn.Conv2d
    bn: Optional[str]      # state_dict prefix of the BatchNorm2d that follows (None: no BN)
    cin: int
    cout: int
    k: int
    stride: int = 1
    bias: bool = False
    relu: bool = False     # ReLU directly after (BN(conv))


@dataclass
class Linear:
    w: str
    fin: int
    fout: int


@dataclass
class Block:
    """BasicBlock (2 convs) or Bottleneck (3 convs); residual add + ReLU after the last conv."""
    convs: List[Conv]
    downsample: Optional[Conv] = None


@dataclass
class FuseTerm:
    kind: str                       # 'id' | 'up' | 'down'
    src: int                        # source branch j
    convs: List[Conv] = field(default_factory=list)
    scale: int = 1                  # nearest-upsample factor for 'up'


@dataclass
class HRModule:
    prefix: str
    branches: List[List[Block]]
    fuse: List[List[FuseTerm]]      # fuse[i] = terms summed into output branch i


@dataclass
class DAModule:
    """One CoAM dual-attention block on one branch (reference DAModule.forward :705-724)."""
    prefix: str
    c: int                          # d_model = branch channels
    cc: int                         # condition channels
    h: int
    w: int
    pos_cnn: Conv
    pos_cnn_cond: Conv
    fc_q: Linear
    fc_k: Linear
    fc_v: Linear
    fc_o: Linear
    chan_cnn: Conv
    chan_cnn_cond: Conv
    chan_fc_o: Linear               # HW -> HW


def _cfg_get(node, key, default=None):
    try:
        return node[key]
    except (KeyError, TypeError):
        return default


class HRNetArch:
    def __init__(self, cfg):
        model = cfg["MODEL"]
        extra = model["EXTRA"]
        self.name = model["NAME"]
        self.num_joints = int(model["NUM_JOINTS"])
        self.img_w, self.img_h = int(model["IMAGE_SIZE"][0]), int(model["IMAGE_SIZE"][1])
        self.use_prenet = bool(_cfg_get(extra, "USE_PRE_NET", False)) and self.name == "pose_hrnet"
        self.use_attention = bool(_cfg_get(extra, "USE_ATTENTION", False)) and self.name == "pose_hrnet_coam"
        ds = cfg["DATASET"]
        stacked = bool(_cfg_get(ds, "STACKED_CONDITION", False))
        colored = bool(_cfg_get(ds, "COLORED", False))
        self.cond_stacked, self.cond_colored = stacked, colored
        final_k = int(_cfg_get(extra, "FINAL_CONV_KERNEL", 1))

        att_modules = list(_cfg_get(model, "ATT_MODULES", [False] * 4))
        selfatt = list(_cfg_get(model, "SELFATT_MODULES", [False] * 4))
        if self.use_attention or self.name == "pose_hrnet_coam":
            for a, s in zip(att_modules, selfatt):
                assert not a or not s           # pose_hrnet_coam.py:354-357
            if any(selfatt):
                raise NotImplementedError("SELFATT_MODULES is never enabled by the reference scripts (SURVEY.md 2 row 1)")
        self.att_modules = att_modules if self.name == "pose_hrnet_coam" else [False] * 4
        self.att_channel_only = bool(_cfg_get(model, "ATT_CHANNEL_ONLY", False))
        self.att_heads = int(_cfg_get(model, "ATTENTION_HEADS", 1))

        # condition channel count seen by CoAM (CoAMBlock.__init__ :736-741)
        self.d_cond = self.num_joints if stacked else (3 if colored else 1)
        # channels of the condition as it arrives in x[:, 3:]
        if self.use_prenet:
            self.in_cond = 3
        elif self.use_attention:
            self.in_cond = self.num_joints if stacked else 3
        else:
            self.in_cond = 0

        # ---- pre-net --------------------------------------------------------------------
        self.prenet_rgb: List[Conv] = []
        self.prenet_cond: List[Conv] = []
        if self.use_prenet:
            self.prenet_rgb = [Conv("rgb_preNet.0", "rgb_preNet.1", 3, 64, 3, 1, bias=True),
                               Conv("rgb_preNet.2", "rgb_preNet.3", 64, 3, 7, 1, bias=True)]
            self.prenet_cond = [Conv("cond_preNet.0", "cond_preNet.1", 3, 3, 7, 1, bias=True)]

        # ---- stem + layer1 --------------------------------------------------------------
        self.stem = [Conv("conv1", "bn1", 3, 64, 3, 2, relu=True),
                     Conv("conv2", "bn2", 64, 64, 3, 2, relu=True)]
        self.layer1: List[Block] = []
        inpl = 64
        for i in range(4):
            p = "layer1.%d" % i
            ds_conv = None
            if inpl != 256:
                ds_conv = Conv(p + ".downsample.0", p + ".downsample.1", inpl, 256, 1, 1)
            self.layer1.append(Block([Conv(p + ".conv1", p + ".bn1", inpl, 64, 1, 1, relu=True),
                                      Conv(p + ".conv2", p + ".bn2", 64, 64, 3, 1, relu=True),
                                      Conv(p + ".conv3", p + ".bn3", 64, 256, 1, 1)], ds_conv))
            inpl = 256

        # ---- transitions + stages -------------------------------------------------------
        self.transitions: List[List[Optional[List[Conv]]]] = []
        self.stages: List[List[HRModule]] = []
        self.stage_channels: List[List[int]] = []
        pre = [256]
        for sidx in (2, 3, 4):
            scfg = extra["STAGE%d" % sidx]
            if scfg["BLOCK"] != "BASIC":
                raise NotImplementedError("only BASIC stage blocks are used by the shipped configs")
            if _cfg_get(scfg, "FUSE_METHOD", "SUM") != "SUM":
                raise NotImplementedError("FUSE_METHOD must be SUM")
            nb = int(scfg["NUM_BRANCHES"])
            chans = [int(c) for c in scfg["NUM_CHANNELS"]]
            nblocks = [int(n) for n in scfg["NUM_BLOCKS"]]
            if nb != len(nblocks):
                raise ValueError("NUM_BRANCHES({}) <> NUM_BLOCKS({})".format(nb, len(nblocks)))
            if nb != len(chans):
                raise ValueError("NUM_BRANCHES({}) <> NUM_CHANNELS({})".format(nb, len(chans)))
            tname = "transition%d" % (sidx - 1)
            trans: List[Optional[List[Conv]]] = []
            for i in range(nb):
                if i < len(pre):
                    if chans[i] != pre[i]:
                        trans.append([Conv("%s.%d.0" % (tname, i), "%s.%d.1" % (tname, i),
                                           pre[i], chans[i], 3, 1, relu=True)])
                    else:
                        trans.append(None)
                else:
                    chain = []
                    hops = i + 1 - len(pre)
                    for j in range(hops):
                        cin = pre[-1]
                        cout = chans[i] if j == hops - 1 else cin
                        chain.append(Conv("%s.%d.%d.0" % (tname, i, j), "%s.%d.%d.1" % (tname, i, j),
                                          cin, cout, 3, 2, relu=True))
                    trans.append(chain)
            self.transitions.append(trans)

            mods: List[HRModule] = []
            nmod = int(scfg["NUM_MODULES"])
            for m in range(nmod):
                multi = not (sidx == 4 and m == nmod - 1)
                mp = "stage%d.%d" % (sidx, m)
                branches = []
                for b in range(nb):
                    blocks = []
                    for k in range(nblocks[b]):
                        bp = "%s.branches.%d.%d" % (mp, b, k)
                        blocks.append(Block([Conv(bp + ".conv1", bp + ".bn1", chans[b], chans[b], 3, 1, relu=True),
                                             Conv(bp + ".conv2", bp + ".bn2", chans[b], chans[b], 3, 1)]))
                    branches.append(blocks)
                fuse: List[List[FuseTerm]] = []
                if nb > 1:
                    for i in range(nb if multi else 1):
                        terms = []
                        for j in range(nb):
                            fp = "%s.fuse_layers.%d.%d" % (mp, i, j)
                            if j == i:
                                terms.append(FuseTerm("id", j))
                            elif j > i:
                                terms.append(FuseTerm("up", j, [Conv(fp + ".0", fp + ".1", chans[j], chans[i], 1, 1)],
                                                      scale=2 ** (j - i)))
                            else:
                                chain = []
                                for k in range(i - j):
                                    last = k == i - j - 1
                                    chain.append(Conv("%s.%d.0" % (fp, k), "%s.%d.1" % (fp, k), chans[j],
                                                      chans[i] if last else chans[j], 3, 2, relu=not last))
                                terms.append(FuseTerm("down", j, chain))
                        fuse.append(terms)
                mods.append(HRModule(mp, branches, fuse))
            self.stages.append(mods)
            self.stage_channels.append(chans)
            pre = chans

        self.final = Conv("final_layer", None, pre[0], self.num_joints, final_k, 1, bias=True)

        # ---- CoAM -----------------------------------------------------------------------
        # att_config[s] places a CoAMBlock in front of stage s+2 (pose_hrnet_coam.py:521,535,554)
        # and att_config[3] after stage4 on branch 0 (:561).  Only [F,T,F,F] is exercised by any
        # shipped script; the other placements are built from the same description.
        # The reference constructs the blocks from ATT_MODULES even when USE_ATTENTION is off
        # (their parameters are in the state_dict); forward only uses them when it is on.
        self.coam: dict = {}
        if self.name == "pose_hrnet_coam" and any(self.att_modules):
            if self.att_heads != 1:
                raise NotImplementedError("ATTENTION_HEADS != 1 is never used by the reference scripts")
            names = ["stage1_att", "stage2_att", "stage3_att", "stage4_att"]
            for s, on in enumerate(self.att_modules):
                if not on:
                    continue
                chans = self.stage_channels[min(s, 2)]
                nbr = len(chans) if s < 3 else 1
                mods = []
                for i in range(nbr):
                    h, w = self.img_h // (4 << i), self.img_w // (4 << i)
                    mods.append(self._da(names[s] + ".att_layers.%d" % i, chans[i], self.d_cond, h, w))
                self.coam[s] = mods

    def _da(self, p: str, c: int, cc: int, h: int, w: int) -> DAModule:
        pp, cp = p + ".position_attention_module", p + ".channel_attention_module"
        return DAModule(
            p, c, cc, h, w,
            pos_cnn=Conv(pp + ".cnn", None, c, c, 3, 1, bias=True),
            pos_cnn_cond=Conv(pp + ".cnn_cond", None, cc, cc, 3, 1, bias=True),
            fc_q=Linear(pp + ".pa.fc_q", cc, c), fc_k=Linear(pp + ".pa.fc_k", c, c),
            fc_v=Linear(pp + ".pa.fc_v", c, c), fc_o=Linear(pp + ".pa.fc_o", c, c),
            chan_cnn=Conv(cp + ".cnn", None, c, c, 3, 1, bias=True),
            chan_cnn_cond=Conv(cp + ".cnn_cond", None, cc, c, 3, 1, bias=True),
            chan_fc_o=Linear(cp + ".pa.fc_o", h * w, h * w))

    # ---- enumeration ---------------------------------------------------------------------
    def convs(self) -> Iterator[Conv]:
        yield from self.prenet_rgb
        yield from self.prenet_cond
        yield from self.stem
        for blk in self.layer1:
            yield from blk.convs
            if blk.downsample:
                yield blk.downsample
        for trans, mods in zip(self.transitions, self.stages):
            for t in trans:
                if t:
                    yield from t
            for m in mods:
                for br in m.branches:
                    for blk in br:
                        yield from blk.convs
                for terms in m.fuse:
                    for t in terms:
                        yield from t.convs
        yield self.final
        for s in sorted(self.coam):
            for da in self.coam[s]:
                if not self.att_channel_only:
                    yield da.pos_cnn
                    yield da.pos_cnn_cond
                yield da.chan_cnn
                yield da.chan_cnn_cond

    def linears(self) -> Iterator[Linear]:
        for s in sorted(self.coam):
            for da in self.coam[s]:
                if not self.att_channel_only:
                    yield from (da.fc_q, da.fc_k, da.fc_v, da.fc_o)
                yield da.chan_fc_o

    def param_specs(self) -> List[Tuple[str, Tuple[int, ...], str]]:
        """(state_dict key, shape, kind); kind in conv_w conv_b bn_w bn_b bn_rm bn_rv bn_nbt lin_w lin_b."""
        out: List[Tuple[str, Tuple[int, ...], str]] = []
        for c in self.convs():
            out.append((c.w + ".weight", (c.cout, c.cin, c.k, c.k), "conv_w"))
            if c.bias:
                out.append((c.w + ".bias", (c.cout,), "conv_b"))
            if c.bn:
                out.append((c.bn + ".weight", (c.cout,), "bn_w"))
                out.append((c.bn + ".bias", (c.cout,), "bn_b"))
                out.append((c.bn + ".running_mean", (c.cout,), "bn_rm"))
                out.append((c.bn + ".running_var", (c.cout,), "bn_rv"))
                out.append((c.bn + ".num_batches_tracked", (), "bn_nbt"))
        for l in self.linears():
            out.append((l.w + ".weight", (l.fout, l.fin), "lin_w"))
            out.append((l.w + ".bias", (l.fout,), "lin_b"))
        return out
