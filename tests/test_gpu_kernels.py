"""-m gpu: every CUDA kernel of the path against a plain torch fp32 reference of the same op on
the same 16-bit-rounded operands (the oracle for whole-network parity is in test_gpu_model.py).
All calls go through the C ABI (ctypes -> libbuctd_b200.so)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from buctd_b200 import _lib as L

from gpu_utils import (TDT, assert_close, ctx, from_act16, pack_w, pads_are_zero, q16, rup, stream, tapgemm, to_act16)

pytestmark = pytest.mark.gpu
DTS = [L.F16, L.BF16]
TOL = {L.F16: 3e-3, L.BF16: 2e-2}      # output rounding of the 16-bit type (2^-11 / 2^-8) + fp32 accumulation order


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cin,cout,h,w,relu,res", [
    (48, 48, 24, 18, True, True), (96, 96, 12, 9, True, False), (192, 192, 12, 9, False, True),
    (384, 384, 12, 9, True, True), (256, 48, 24, 18, True, False), (64, 64, 10, 7, True, False),
    (17, 48, 24, 18, False, False), (17, 17, 12, 9, False, False), (3, 96, 8, 6, False, False),
])
def test_tapgemm_conv3x3_flat(dt, cin, cout, h, w, relu, res):
    B = 3
    x = q16(_rand(B, cin, h, w, seed=1), dt)
    wt = q16(_rand(cout, cin, 3, 3, seed=2, scale=1.0 / math.sqrt(9 * cin)), dt)
    bias = _rand(cout, seed=3, scale=0.1)
    r = q16(_rand(B, cout, h, w, seed=4), dt) if res else None
    a16 = to_act16(x, dt, cs=rup(cin, 16) if cin % 8 else None)
    out = torch.full((B, h + 1, w + 1, rup(cout, 16)), 7.0, dtype=TDT[dt], device="cuda")
    r16 = to_act16(r, dt, cs=out.shape[-1]) if res else None
    tapgemm(dt, a16, wt, bias, out, H=h, W=w, relu=relu, residual=r16, k=cin)
    ref = F.conv2d(x.cpu(), wt.cpu(), bias.cpu(), padding=1)
    if res:
        ref = ref + r.cpu()
    if relu:
        ref = F.relu(ref)
    assert_close(from_act16(out, cout), ref, TOL[dt], "conv3x3 %d->%d" % (cin, cout))
    assert pads_are_zero(out)
    if out.shape[-1] > cout:
        assert out[..., cout:].float().abs().max() == 0


@pytest.mark.parametrize("msub", [1, 2, 3, 4])
@pytest.mark.parametrize("cin,cout", [(48, 48), (96, 96), (64, 64), (256, 48), (192, 192)])
def test_tapgemm_slab_supertiles(monkeypatch, msub, cin, cout):
    """slab kernel: every super-tile size (row-shifted tap views across sub-tiles), resident and ring weights"""
    monkeypatch.setenv("BUCTD_FORCE_MSUB", str(msub))
    dt, B, h, w = L.F16, 7, 24, 18
    x = q16(_rand(B, cin, h, w, seed=41), dt)
    wt = q16(_rand(cout, cin, 3, 3, seed=42, scale=1.0 / math.sqrt(9 * cin)), dt)
    bias = _rand(cout, seed=43, scale=0.1)
    r = q16(_rand(B, cout, h, w, seed=44), dt)
    out = torch.full((B, h + 1, w + 1, cout), 7.0, dtype=TDT[dt], device="cuda")
    tapgemm(dt, to_act16(x, dt), wt, bias, out, H=h, W=w, relu=True, residual=to_act16(r, dt))
    ref = F.relu(F.conv2d(x.cpu(), wt.cpu(), bias.cpu(), padding=1) + r.cpu())
    assert_close(from_act16(out, cout), ref, TOL[dt], "slab msub=%d %d->%d" % (msub, cin, cout))
    assert pads_are_zero(out)


@pytest.mark.parametrize("cin,cout,B,h,w", [(48, 48, 10, 96, 72), (96, 96, 24, 48, 36), (384, 384, 40, 12, 9)])
def test_tapgemm_slab_persistent_many_items(cin, cout, B, h, w):
    """more work items than SMs: every CTA loops, both TMEM accumulator sets and the slab ring wrap"""
    dt = L.F16
    x = q16(_rand(B, cin, h, w, seed=45), dt)
    wt = q16(_rand(cout, cin, 3, 3, seed=46, scale=1.0 / math.sqrt(9 * cin)), dt)
    bias = _rand(cout, seed=47, scale=0.1)
    out = torch.zeros(B, h + 1, w + 1, cout, dtype=TDT[dt], device="cuda")
    tapgemm(dt, to_act16(x, dt), wt, bias, out, H=h, W=w, relu=True)
    ref = F.relu(F.conv2d(x.cpu(), wt.cpu(), bias.cpu(), padding=1))
    assert_close(from_act16(out, cout), ref, TOL[dt], "slab persistent")
    assert pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cin,cout,h,w", [(64, 256, 12, 9), (256, 64, 12, 9), (96, 48, 6, 4), (384, 192, 3, 2), (48, 96, 24, 18)])
def test_tapgemm_conv1x1(dt, cin, cout, h, w):
    B = 2
    x = q16(_rand(B, cin, h, w, seed=5), dt)
    wt = q16(_rand(cout, cin, 1, 1, seed=6, scale=1.0 / math.sqrt(cin)), dt)
    bias = _rand(cout, seed=7, scale=0.1)
    out = torch.zeros(B, h + 1, w + 1, cout, dtype=TDT[dt], device="cuda")
    tapgemm(dt, to_act16(x, dt), wt, bias, out, H=h, W=w)
    assert_close(from_act16(out, cout), F.conv2d(x.cpu(), wt.cpu(), bias.cpu()), TOL[dt], "conv1x1")
    assert pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cin,cout,h,w,relu", [(48, 96, 24, 18, False), (64, 64, 48, 36, True), (96, 192, 12, 10, True),
                                               (192, 384, 24, 18, True), (48, 48, 96, 72, True), (256, 96, 24, 18, True)])
def test_tapgemm_conv3x3_stride2_patch(dt, cin, cout, h, w, relu):
    B = 2
    x = q16(_rand(B, cin, h, w, seed=8), dt)
    wt = q16(_rand(cout, cin, 3, 3, seed=9, scale=1.0 / math.sqrt(9 * cin)), dt)
    bias = _rand(cout, seed=10, scale=0.1)
    oh, ow = h // 2, w // 2
    out = torch.zeros(B, oh + 1, ow + 1, cout, dtype=TDT[dt], device="cuda")
    tapgemm(dt, to_act16(x, dt), wt, bias, out, H=oh, W=ow, stride=2, relu=relu, in_hw=(h, w))
    ref = F.conv2d(x.cpu(), wt.cpu(), bias.cpu(), stride=2, padding=1)
    if relu:
        ref = F.relu(ref)
    assert_close(from_act16(out, cout), ref, TOL[dt], "conv3x3s2 %d->%d" % (cin, cout))
    assert pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
def test_tapgemm_head_nchw32(dt):
    B, cin, K, h, w = 2, 48, 17, 24, 18
    x = q16(_rand(B, cin, h, w, seed=11), dt)
    wt = q16(_rand(K, cin, 1, 1, seed=12, scale=0.2), dt)
    bias = _rand(K, seed=13, scale=0.1)
    out = torch.full((B, K, h, w), 9.0, dtype=torch.float32, device="cuda")
    tapgemm(dt, to_act16(x, dt), wt, bias, out, H=h, W=w, out_mode=L.OUT_NCHW32)
    assert_close(out, F.conv2d(x.cpu(), wt.cpu(), bias.cpu()), 1e-4, "head")   # fp32 output: no 16-bit rounding


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,cc,H,W", [(2, 17, 128, 96), (3, 3, 384, 288), (1, 1, 256, 192), (2, 20, 96, 64)])
def test_cond_resize_pyramid(dt, B, cc, H, W):
    """the fused three-scale resize == three single-scale launches (bit for bit) == torch antialiased bilinear"""
    x = (_rand(B, 3 + cc, H, W, seed=25).abs() * 100).contiguous()
    cs = rup(cc, 16)
    outs = [torch.zeros(B, H // s + 1, W // s + 1, cs, dtype=TDT[dt], device="cuda") for s in (4, 8, 16)]
    a = L.Resize3Args()
    a.dtype, a.batch, a.x_channels, a.c_off, a.cc, a.h, a.w = dt, B, 3 + cc, 3, cc, H, W
    for i in range(3):
        a.out_cs[i] = cs
        a.out[i] = outs[i].data_ptr()
    a.x = x.data_ptr()
    ctx().call("buctd_cond_resize3", a, stream())
    torch.cuda.synchronize()
    for i, s in enumerate((4, 8, 16)):
        oh, ow = H // s, W // s
        single = torch.zeros_like(outs[i])
        r = L.ResizeArgs()
        r.dtype, r.batch, r.x_channels, r.c_off, r.cc = dt, B, 3 + cc, 3, cc
        r.h, r.w, r.oh, r.ow, r.out_cs, r.antialias = H, W, oh, ow, cs, 1
        r.x, r.out = x.data_ptr(), single.data_ptr()
        ctx().run(L.OP_RESIZE, r, stream())
        torch.cuda.synchronize()
        ref = F.interpolate(x[:, 3:].cpu(), size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)
        assert_close(from_act16(outs[i], cc), ref, TOL[dt], "pyramid x%d" % s)
        assert torch.equal(outs[i], single), "pyramid scale %d differs from the single-scale kernel" % s
        assert pads_are_zero(outs[i])


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("rows,k,n", [(96, 432, 432), (200, 768, 768), (130, 1728, 1728)])
def test_tapgemm_linear_rows(dt, rows, k, n):
    x = q16(_rand(rows, k, seed=14), dt)
    wt = q16(_rand(n, k, seed=15, scale=1.0 / math.sqrt(k)), dt)
    bias = _rand(n, seed=16, scale=0.1)
    a16 = x.to(TDT[dt]).contiguous()
    out = torch.zeros(rows, n, dtype=TDT[dt], device="cuda")
    tapgemm(dt, a16, wt.view(n, k, 1, 1), bias, out, H=0, W=0, out_mode=L.OUT_ROWS16, m_rows=rows, batch=1)
    assert_close(out, F.linear(x.cpu(), wt.cpu(), bias.cpu()), TOL[dt], "linear")


@pytest.mark.parametrize("dt", DTS)
def test_stem_conv(dt):
    B, h, w = 2, 32, 24
    x = _rand(B, 5, h, w, seed=17)
    wt = _rand(64, 3, 3, 3, seed=18, scale=0.3)
    bias = _rand(64, seed=19, scale=0.1)
    out = torch.zeros(B, h // 2 + 1, w // 2 + 1, 64, dtype=TDT[dt], device="cuda")
    a = L.StemArgs()
    a.dtype, a.batch, a.x_channels, a.h, a.w, a.out_cs = dt, B, 5, h, w, 64
    wf = wt.reshape(64, 27).contiguous()
    a.x, a.weight, a.bias, a.out = x.data_ptr(), wf.data_ptr(), bias.data_ptr(), out.data_ptr()
    ctx().run(L.OP_STEM, a, stream())
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x[:, :3].cpu(), wt.cpu(), bias.cpu(), stride=2, padding=1))
    assert_close(from_act16(out, 64), ref, TOL[dt], "stem")
    assert pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
def test_hr_fuse(dt):
    B, C, h, w = 2, 48, 24, 16
    t0 = q16(_rand(B, C, h, w, seed=20), dt)
    t1 = q16(_rand(B, C, h // 2, w // 2, seed=21), dt)
    t2 = q16(_rand(B, C, h // 4, w // 4, seed=22), dt)
    t3 = q16(_rand(B, C, h // 8, w // 8, seed=23), dt)
    terms = [to_act16(t, dt) for t in (t0, t1, t2, t3)]
    out = torch.full((B, h + 1, w + 1, C), 3.0, dtype=TDT[dt], device="cuda")
    a = L.FuseArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.out_cs, a.nterms, a.relu = dt, B, h, w, C, C, 4, 1
    for j, t in enumerate(terms):
        a.term[j], a.term_cs[j], a.term_shift[j] = t.data_ptr(), C, j
    a.out = out.data_ptr()
    ctx().run(L.OP_FUSE, a, stream())
    torch.cuda.synchronize()
    ref = t0 + sum(F.interpolate(t, scale_factor=2 ** (j + 1), mode="nearest") for j, t in enumerate((t1, t2, t3)))
    assert_close(from_act16(out, C), F.relu(ref), TOL[dt], "fuse")
    assert pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("aa", [1, 0])
def test_cond_resize(dt, aa):
    B, cc, H, W = 2, 17, 128, 96
    x = _rand(B, 3 + cc, H, W, seed=24).abs() * 100
    for s in (4, 8, 16):
        oh, ow = H // s, W // s
        out = torch.zeros(B, oh + 1, ow + 1, 32, dtype=TDT[dt], device="cuda")
        a = L.ResizeArgs()
        a.dtype, a.batch, a.x_channels, a.c_off, a.cc = dt, B, 3 + cc, 3, cc
        a.h, a.w, a.oh, a.ow, a.out_cs, a.antialias = H, W, oh, ow, 32, aa
        a.x, a.out = x.data_ptr(), out.data_ptr()
        ctx().run(L.OP_RESIZE, a, stream())
        torch.cuda.synchronize()
        ref = F.interpolate(x[:, 3:].cpu(), size=(oh, ow), mode="bilinear", align_corners=False, antialias=bool(aa))
        assert_close(from_act16(out, cc), ref, TOL[dt], "resize x%d aa=%d" % (s, aa))
        assert out[..., cc:].float().abs().max() == 0 and pads_are_zero(out)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("C,h,w", [(48, 24, 18), (96, 12, 9), (192, 6, 5), (32, 8, 6), (64, 4, 3), (384, 12, 9), (256, 4, 3)])
def test_channel_attention(dt, C, h, w):
    B = 2
    y = q16(_rand(B, C, h, w, seed=25), dt)
    yc = q16(_rand(B, C, h, w, seed=26), dt)
    out = torch.zeros(B, C, h * w, dtype=TDT[dt], device="cuda")
    nch = L.chan_att_scratch_floats(B, h, w, C) // (B * C * C)
    scratch = torch.zeros(B, nch, C, C, dtype=torch.float32, device="cuda")
    ya, yca = to_act16(y, dt), to_act16(yc, dt)
    a = L.ChanAttArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.y_cs, a.yc_cs = dt, B, h, w, C, C, C
    a.y, a.yc, a.scratch, a.out = ya.data_ptr(), yca.data_ptr(), scratch.data_ptr(), out.data_ptr()
    ctx().run(L.OP_CHAN_ATT, a, stream())
    torch.cuda.synchronize()
    yf, ycf = y.view(B, C, -1), yc.view(B, C, -1)
    att = torch.softmax(torch.matmul(ycf, yf.transpose(1, 2)) / math.sqrt(h * w), -1)
    assert_close(scratch[:, 0], att, 1e-4, "chan att probs")
    assert_close(out, torch.matmul(att, yf), TOL[dt], "chan att out")
    # padded pixel order (what the plan uses): out[b][c1][y (w + 1) + x], tensor-core apply when there are >= 2 scratch chunks
    # (probabilities rounded to 16 bit before the product), CUDA-core kernels otherwise -- same layout either way
    rows1 = (h + 1) * (w + 1)
    ld = (rows1 + 7) // 8 * 8
    outp = torch.zeros(B, C, ld, dtype=TDT[dt], device="cuda")
    scratch.zero_()
    a.out, a.out_stride, a.out_padded = outp.data_ptr(), ld, 1
    ctx().run(L.OP_CHAN_ATT, a, stream())
    torch.cuda.synchronize()
    assert_close(scratch[:, 0], att, 1e-4, "chan att probs (padded run)")
    got = outp[:, :, :rows1].view(B, C, h + 1, w + 1)
    assert got[:, :, h].float().abs().max() == 0 and got[:, :, :, w].float().abs().max() == 0
    assert_close(got[:, :, :h, :w].reshape(B, C, h * w), torch.matmul(att, yf), 2 * TOL[dt], "chan att out (padded order)")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("C,h,w", [(48, 24, 18), (96, 12, 9), (192, 7, 5), (32, 16, 12), (64, 8, 6), (128, 4, 3),
                                   # every head dim runs on tcgen05: channel blocks of 64 (1..6), several key tiles, with and
                                   # without the ones column (C % 64 != 0 and <= 2 blocks)
                                   (96, 48, 36), (192, 24, 18), (128, 16, 12), (256, 4, 3), (384, 12, 9), (16, 9, 7), (112, 20, 13),
                                   # two query tiles per CTA (C <= 64, >= 2 tiles): odd tile count, no ones column (C = 64)
                                   (48, 28, 20), (64, 24, 18), (48, 48, 36)])
def test_position_attention(dt, C, h, w):
    B = 2
    q = q16(_rand(B, C, h, w, seed=27, scale=1.5), dt)
    k = q16(_rand(B, C, h, w, seed=28), dt)
    v = q16(_rand(B, C, h, w, seed=29), dt)
    qa = to_act16(q, dt)
    kv = to_act16(torch.cat([k, v], 1), dt)
    # poison the pad rows of k/v: they must be masked out, not merely zero
    kv[:, -1] = 50.0
    kv[:, :, -1] = 50.0
    o = torch.zeros(B, h + 1, w + 1, C, dtype=TDT[dt], device="cuda")
    a = L.PosAttArgs()
    a.dtype, a.batch, a.h, a.w, a.c = dt, B, h, w, C
    a.q_cs, a.k_cs, a.v_cs, a.o_cs = C, 2 * C, 2 * C, C
    a.q, a.k, a.v, a.o = qa.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * C, o.data_ptr()
    ctx().run(L.OP_POS_ATT, a, stream())
    torch.cuda.synchronize()
    qt, kt, vt = [t.view(B, C, -1).transpose(1, 2) for t in (q, k, v)]
    att = torch.softmax(torch.matmul(qt, kt.transpose(1, 2)) / math.sqrt(C), -1)
    ref = torch.matmul(att, vt).transpose(1, 2).reshape(B, C, h, w)
    assert_close(from_act16(o, C), ref, 2.5 * TOL[dt], "pos att")     # P is rounded to 16 bit before P.V
    assert pads_are_zero(o)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("channel_only", [0, 1])
def test_da_combine(dt, channel_only):
    B, C, h, w = 2, 48, 11, 7
    x = q16(_rand(B, C, h, w, seed=30), dt)
    p = q16(_rand(B, C, h, w, seed=31), dt)
    c = q16(_rand(B, C, h, w, seed=32), dt)
    out = torch.zeros(B, h + 1, w + 1, C, dtype=TDT[dt], device="cuda")
    xa, pa, cp = to_act16(x, dt), to_act16(p, dt), c.view(B, C, -1).to(TDT[dt]).contiguous()
    a = L.DaCombineArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.x_cs, a.p_cs, a.out_cs, a.channel_only = dt, B, h, w, C, C, C, C, channel_only
    a.x, a.p_out, a.c_out, a.out = xa.data_ptr(), pa.data_ptr(), cp.data_ptr(), out.data_ptr()
    ctx().run(L.OP_DA_COMBINE, a, stream())
    torch.cuda.synchronize()
    ref = x * c if channel_only else x + (p + c)
    assert_close(from_act16(out, C), ref, TOL[dt], "da_combine")
    assert pads_are_zero(out)


def test_prenet_composed():
    """composed 9x9 + border variants == conv7x7(BN(conv3x3(rgb))) + conv7x7(cond) in fp32"""
    from buctd_b200 import pose_hrnet
    from buctd_b200.config import make_cfg
    from buctd_b200.synth import synth_state_dict
    cfg = make_cfg("prenet", 32, (24, 40), cond="colored")
    net = pose_hrnet.get_pose_net(cfg, is_train=False)
    sd = synth_state_dict(net.arch, 3)
    net.load_state_dict(sd)
    f = {k: v.cuda() for k, v in net._compose_prenet().items()}
    B, H, W = 2, 40, 24
    x = _rand(B, 6, H, W, seed=33)
    out = torch.zeros(B, 3, H, W, device="cuda")
    a = L.PrenetArgs()
    a.batch, a.x_channels, a.h, a.w = B, 6, H, W
    a.x, a.w_eff, a.b_eff, a.wc, a.out = x.data_ptr(), f["prenet.w_eff"].data_ptr(), f["prenet.b_eff"].data_ptr(), f["prenet.wc"].data_ptr(), out.data_ptr()
    ctx().run(L.OP_PRENET, a, stream())
    torch.cuda.synchronize()
    sdc = {k: v.cuda() for k, v in sd.items()}

    def bn(p, t):
        return F.batch_norm(t, sdc[p + ".running_mean"], sdc[p + ".running_var"], sdc[p + ".weight"], sdc[p + ".bias"], False, 0.0, 1e-5)
    t = bn("rgb_preNet.1", F.conv2d(x[:, :3], sdc["rgb_preNet.0.weight"], sdc["rgb_preNet.0.bias"], padding=1))
    x0 = bn("rgb_preNet.3", F.conv2d(t, sdc["rgb_preNet.2.weight"], sdc["rgb_preNet.2.bias"], padding=3))
    x1 = bn("cond_preNet.1", F.conv2d(x[:, 3:], sdc["cond_preNet.0.weight"], sdc["cond_preNet.0.bias"], padding=3))
    assert_close(out, x0 + x1, 2e-5, "prenet")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("c,B,h,w,nblocks,dxf", [
    (48, 7, 24, 18, 2, "0"), (96, 5, 12, 9, 4, "1"), (192, 9, 12, 9, 2, "1"), (48, 40, 48, 36, 4, "0"),
    # dx-fused plan (C <= 64: one N = 3C MMA per kernel row, 126-row tiles), default and pinned super-tile sizes
    (48, 7, 24, 18, 2, "1"), (48, 40, 48, 36, 4, "1"), (48, 40, 48, 36, 4, "1:2"), (48, 3, 5, 4, 1, "1"), (48, 1, 96, 72, 1, "1:3"),
    # > kLag + 1 items per CTA and layer: the publish FIFO runs across layer boundaries (no drain)
    (48, 16, 96, 72, 2, "1"), (48, 40, 96, 72, 2, "0"),
    (192, 30, 24, 18, 4, "0"), (96, 33, 48, 36, 3, "0"),        # few items per CTA and layer (drain at every layer end), partial last round
    (32, 9, 32, 24, 2, "1"), (32, 9, 32, 24, 2, "1:4"), (64, 6, 20, 15, 2, "2"), (64, 6, 20, 15, 2, "2:2")])
def test_tapgemm_chain_basic_blocks(dt, c, B, h, w, nblocks, dxf, monkeypatch):
    """chain launch: nblocks BasicBlocks (conv-relu-conv-add-relu) in ONE persistent kernel with grid barriers,
    against the same sequence in torch on 16-bit-rounded operands"""
    monkeypatch.setenv("BUCTD_DXF", dxf.split(":")[0])
    if ":" in dxf:
        monkeypatch.setenv("BUCTD_DXF_MSUB", dxf.split(":")[1])
    else:
        monkeypatch.delenv("BUCTD_DXF_MSUB", raising=False)
    _chain_case(dt, c, B, h, w, nblocks)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("pair", ["2", "0"])
@pytest.mark.parametrize("c,B,h,w,nblocks", [
    (96, 5, 12, 9, 4), (192, 30, 24, 18, 4), (96, 33, 48, 36, 3),
    (192, 9, 12, 9, 2),            # 10 tiles: 5 pairs
    (192, 7, 12, 9, 2),            # 910 rows = 8 tiles (7.1): last tile partial
    (384, 5, 12, 9, 2),            # two N tiles of 192; 650 rows = 6 tiles (5.1)
    (384, 9, 12, 9, 1),            # 1170 rows = 10 tiles x 2 N tiles
    (384, 64, 12, 9, 4),           # branch 3 of W48 at 384x288, B = 64: 65 tiles -> 33 units, the last one half dummy
    (128, 3, 24, 18, 2), (256, 11, 12, 9, 2)])
def test_tapgemm_chain_pair(dt, pair, c, B, h, w, nblocks, monkeypatch):
    """the wide chains as CTA pairs (cta_group::2, half weight tiles per CTA; odd tile counts leave a dummy tile) and,
    with BUCTD_CHAIN_PAIR=0, as the one-CTA plan (=2: pairs from C = 96 on; the default starts at 192)"""
    monkeypatch.setenv("BUCTD_CHAIN_PAIR", pair)
    _chain_case(dt, c, B, h, w, nblocks)


def _chain_case(dt, c, B, h, w, nblocks):
    x = q16(_rand(B, c, h, w, seed=61), dt)
    ws = [q16(_rand(c, c, 3, 3, seed=62 + i, scale=1.0 / math.sqrt(9 * c)), dt) for i in range(2 * nblocks)]
    bs = [_rand(c, seed=80 + i, scale=0.1).float().contiguous() for i in range(2 * nblocks)]
    bufs = [to_act16(x, dt), torch.zeros(B, h + 1, w + 1, c, dtype=TDT[dt], device="cuda"),
            torch.zeros(B, h + 1, w + 1, c, dtype=TDT[dt], device="cuda")]
    packed = [pack_w(wt, dt) for wt in ws]
    a = L.TapGemmChainArgs()
    a.n_layers = 2 * nblocks
    cur = 0
    ref = x.cpu()
    for blk in range(nblocks):
        t, o = (cur + 1) % 3, (cur + 2) % 3
        for j, (src, dst, res) in enumerate(((cur, t, None), (t, o, cur))):
            li = 2 * blk + j
            g = a.layer[li]
            wp, n_alloc, k_cs = packed[li]
            g.dtype, g.mode, g.ntaps, g.stride, g.batch = dt, L.TG_FLAT, 9, 1, B
            g.out_h, g.out_w, g.in_h, g.in_w = h, w, h, w
            g.m_rows = B * (h + 1) * (w + 1)
            g.k, g.a_cs, g.n, g.n_alloc, g.w_cs = c, c, c, n_alloc, k_cs
            g.out_cs, g.res_cs, g.relu, g.out_mode = c, (c if res is not None else 0), 1, L.OUT_ACT16
            g.a, g.w, g.bias, g.out = bufs[src].data_ptr(), wp.data_ptr(), bs[li].data_ptr(), bufs[dst].data_ptr()
            g.residual = bufs[res].data_ptr() if res is not None else None
        # torch reference with the same 16-bit rounding between layers
        mid = q16(F.relu(F.conv2d(ref, ws[2 * blk].cpu(), bs[2 * blk].cpu(), padding=1)), dt)
        ref = q16(F.relu(F.conv2d(mid, ws[2 * blk + 1].cpu(), bs[2 * blk + 1].cpu(), padding=1) + ref), dt)
        cur = o
    ctx().run(L.OP_CHAIN, a, stream())
    torch.cuda.synchronize()
    assert_close(from_act16(bufs[cur], c), ref, 4 * TOL[dt], "chain of %d blocks C=%d" % (nblocks, c))
    assert pads_are_zero(bufs[cur])


def _linear_rows_case(dt, rows, k, n, relu, seed=40):
    """large nn.Linear on a [rows, k] matrix with padded row strides (the channel attention's fc_o on the planar matrix)"""
    lda, ldo = rup(k, 8), rup(n, 8)
    x = q16(_rand(rows, k, seed=seed), dt)
    wt = q16(_rand(n, k, seed=seed + 1, scale=1.0 / math.sqrt(k)), dt)
    bias = _rand(n, seed=seed + 2, scale=0.1)
    a16 = torch.zeros(rows, lda, dtype=TDT[dt], device="cuda")
    a16[:, :k] = x.to(TDT[dt])
    out = torch.full((rows, ldo), 7.0, dtype=TDT[dt], device="cuda")
    tapgemm(dt, a16, wt.view(n, k, 1, 1), bias, out, H=0, W=0, out_mode=L.OUT_ROWS16, m_rows=rows, batch=1, k=k, relu=relu)
    ref = F.linear(x.cuda(), wt.cuda(), bias.cuda())          # fp32 on the GPU (the big case is 150 GFLOP)
    if relu:
        ref = ref.relu()
    return out, ref


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("rows,k,n,relu", [(512, 481, 432, False), (300, 1825, 1728, True), (256, 256, 256, False),
                                           (1000, 7081, 1000, False)])
def test_pairgemm_linear_rows(dt, rows, k, n, relu, monkeypatch):
    """CTA-pair (cta_group::2) GEMM: M / N / K tails, bias, ReLU, and the same numbers as the one-CTA slab plan"""
    out, ref = _linear_rows_case(dt, rows, k, n, relu)
    assert_close(out[:, :n], ref, TOL[dt], "pairgemm")
    monkeypatch.setenv("BUCTD_PAIRGEMM", "0")
    out1, _ = _linear_rows_case(dt, rows, k, n, relu)
    # both plans accumulate the same fp32 products in the same k order on the tensor core
    assert_close(out[:, :n], out1[:, :n], 2e-3, "pairgemm vs slab plan")


def test_pairgemm_fc_o_w48():
    """the real shape: Linear(6912, 6912) over 64 x 48 planar rows of padded-pixel length 7081"""
    out, ref = _linear_rows_case(L.F16, 3072, 7081, 6912, False)
    assert_close(out[:, :6912], ref, TOL[L.F16], "pairgemm fc_o")
