"""-m gpu: the kernels of the training step (row a-T) one by one through the C ABI, each against a plain PyTorch fp32
reference of the same op on 16-bit-rounded inputs (the autograd of the reference's nn.Conv2d / nn.BatchNorm2d /
nn.Linear / torch.matmul / softmax / nn.Upsample, lib/core/function.py:130-138)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from buctd_b200 import _lib as L
from gpu_utils import TDT, assert_close, ctx, from_act16, pads_are_zero, q16, stream, to_act16

pytestmark = pytest.mark.gpu
DTS = [L.F16, L.BF16]
TOL = {L.F16: 4e-3, L.BF16: 2.5e-2}


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


def drop_mask(seed, crop, rows, cols, p):
    """the hash mask of include/buctd_b200.h ("Dropout"): keep[row][col] for one crop, numpy uint32 arithmetic"""
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(cols, dtype=np.uint64)[None, :]
    M = np.uint64(0xFFFFFFFF)
    x = (np.uint64(seed) ^ ((r * np.uint64(0x9E3779B1) + c * np.uint64(0x85EBCA77) + np.uint64(crop) * np.uint64(0xC2B2AE3D)) & M)) & M
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    thr = np.uint64(min(int(p * 4294967296.0), 4294967295))
    return x >= thr


def gemm16(dt, a, b, m, n, k, *, a_mn, b_mn, nbatch=1, nsplit=1, ntaps=1, tap_off=None, out_f32=True, alpha=1.0, accumulate=False,
           d=None, lda=None, ldb=None, a_bs=0, b_bs=0, b_rows=0, ldd=None, d_strides=(0, 0, 0), tap_group=0):
    g = L.Gemm16Args()
    g.dtype, g.a_mn_major, g.b_mn_major, g.m, g.n, g.k = dt, int(a_mn), int(b_mn), m, n, k
    g.nbatch, g.nsplit, g.ntaps, g.b_rows = nbatch, nsplit, ntaps, b_rows
    g.lda, g.ldb = lda or a.shape[-1], ldb or b.shape[-1]
    g.ldd = ldd or n
    g.out_f32, g.accumulate, g.tap_group = int(out_f32), int(accumulate), tap_group
    g.a_batch_stride, g.b_batch_stride = a_bs, b_bs
    g.d_t_stride, g.d_s_stride, g.d_b_stride = d_strides
    for i, o in enumerate(tap_off or [0] * 9):
        g.tap_off[i] = o
    g.alpha = alpha
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), d.data_ptr()
    ctx().call("buctd_gemm16", g, stream())
    torch.cuda.synchronize()
    return d


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("m,n,k", [(200, 48, 300), (128, 256, 64), (70, 600, 130), (384, 96, 1000)])
def test_gemm16_majors(dt, a_mn, b_mn, m, n, k):
    A = q16(_rand(m, k, seed=1), dt)
    B = q16(_rand(n, k, seed=2), dt)
    rup = lambda v: (v + 7) // 8 * 8
    if a_mn:
        a = torch.zeros(k, rup(m), dtype=TDT[dt], device="cuda"); a[:, :m] = A.t()
    else:
        a = torch.zeros(m, rup(k), dtype=TDT[dt], device="cuda"); a[:, :k] = A
    if b_mn:
        b = torch.zeros(k, rup(n), dtype=TDT[dt], device="cuda"); b[:, :n] = B.t()
    else:
        b = torch.zeros(n, rup(k), dtype=TDT[dt], device="cuda"); b[:, :k] = B
    d = torch.full((m, n), 7.0, device="cuda")
    gemm16(dt, a, b, m, n, k, a_mn=a_mn, b_mn=b_mn, d=d, alpha=0.5)
    assert_close(d, 0.5 * A @ B.t(), 2e-3, "gemm16 f32 out")
    d16 = torch.zeros(m, rup(n), dtype=TDT[dt], device="cuda")
    gemm16(dt, a, b, m, n, k, a_mn=a_mn, b_mn=b_mn, d=d16, out_f32=False, ldd=rup(n))
    assert_close(d16[:, :n], A @ B.t(), TOL[dt], "gemm16 16-bit out")
    d.fill_(1.0)
    gemm16(dt, a, b, m, n, k, a_mn=a_mn, b_mn=b_mn, d=d, accumulate=True)
    assert_close(d, 1.0 + A @ B.t(), 2e-3, "gemm16 accumulate")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("group", [0, 3])
@pytest.mark.parametrize("cin,cout,B,h,w,nsplit", [(48, 48, 3, 12, 9, 4), (96, 64, 2, 7, 5, 3), (32, 200, 2, 6, 4, 1), (17, 48, 2, 8, 6, 2),
                                                   (192, 192, 2, 6, 5, 2), (384, 384, 2, 3, 2, 1), (256, 48, 1, 9, 7, 3)])
def test_gemm16_conv_weight_gradient(dt, cin, cout, B, h, w, nsplit, group):
    """dW[co][ci][ky][kx] of a 3x3 stride-1 convolution: gemm16 (taps = row offsets of the padded act16 matrices, split
    partials) + wgrad_reduce against torch's conv2d weight gradient."""
    x = q16(_rand(B, cin, h, w, seed=3), dt)
    dy = q16(_rand(B, cout, h, w, seed=4), dt)
    xa, dya = to_act16(x, dt), to_act16(dy, dt)
    rows = B * (h + 1) * (w + 1)
    part = torch.full((nsplit, 9, cout, cin), 3.0, device="cuda")
    offs = [(ky - 1) * (w + 1) + (kx - 1) for ky in range(3) for kx in range(3)]
    gemm16(dt, dya, xa, cout, cin, rows, a_mn=1, b_mn=1, nsplit=nsplit, ntaps=9, tap_off=offs, d=part, ldd=cin,
           d_strides=(cout * cin, 9 * cout * cin, 0), b_rows=rows, tap_group=group)
    grad = torch.zeros(cout, cin, 3, 3, device="cuda")
    r = L.WgradReduceArgs()
    r.n, r.k, r.ntaps, r.nsplit, r.accumulate, r.inv_scale = cout, cin, 9, nsplit, 0, 0.25
    r.partial, r.grad = part.data_ptr(), grad.data_ptr()
    ctx().call("buctd_wgrad_reduce", r, stream())
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), dy, padding=1) * 0.25
    assert_close(grad, ref, 3e-3, "conv weight gradient")


@pytest.mark.parametrize("dt", DTS)
def test_gemm16_batched_attention_products(dt):
    """the five products of the attention backward on per-crop matrices (batch = crops, K never crosses a crop)"""
    B, n, C = 3, 150, 48
    ld = 152
    Q, K_, dO = [q16(_rand(B, n, C, seed=s), dt) for s in (5, 6, 7)]
    P = q16(_rand(B, n, n, seed=8).abs(), dt)
    q, k, do = [t.to(TDT[dt]).contiguous() for t in (Q, K_, dO)]
    p = torch.zeros(B, n, ld, dtype=TDT[dt], device="cuda"); p[:, :, :n] = P
    s = torch.zeros(B, n, ld, device="cuda")
    gemm16(dt, q, k, n, n, C, a_mn=0, b_mn=0, nbatch=B, a_bs=n * C, b_bs=n * C, d=s, ldd=ld, d_strides=(0, 0, n * ld), alpha=0.3)
    assert_close(s[:, :, :n], 0.3 * Q @ K_.transpose(1, 2), 2e-3, "S = Q K^T")
    dv = torch.zeros(B, n, C, dtype=TDT[dt], device="cuda")
    gemm16(dt, p, do, n, C, n, a_mn=1, b_mn=1, nbatch=B, a_bs=n * ld, b_bs=n * C, d=dv, out_f32=False, ldd=C, d_strides=(0, 0, n * C), lda=ld)
    assert_close(dv, P.transpose(1, 2) @ dO, TOL[dt], "dV = P^T dO")
    dq = torch.zeros(B, n, C, dtype=TDT[dt], device="cuda")
    gemm16(dt, p, k, n, C, n, a_mn=0, b_mn=1, nbatch=B, a_bs=n * ld, b_bs=n * C, d=dq, out_f32=False, ldd=C, d_strides=(0, 0, n * C), lda=ld)
    assert_close(dq, P @ K_, TOL[dt], "dQ = dS K")


def _bn_case(dt, B=3, C=48, h=9, w=7, seed=10):
    z = q16(_rand(B, C, h, w, seed=seed, scale=2.0) + 0.5, dt)
    return z, to_act16(z, dt)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("C,h,w", [(48, 9, 7), (96, 5, 4), (384, 3, 2), (17, 6, 5), (256, 12, 9)])
def test_bn_train_forward_backward(dt, C, h, w):
    B = 3
    z, za = _bn_case(dt, B, C, h, w)
    cs = za.shape[-1]
    gamma, beta = _rand(C, seed=11) * 0.5 + 1.0, _rand(C, seed=12) * 0.3
    rm, rv = _rand(C, seed=13), _rand(C, seed=14).abs() + 0.5
    res = q16(_rand(B, C, h, w, seed=15), dt)
    resa = to_act16(res, dt)
    mr = torch.zeros(2, C, device="cuda")
    scratch = torch.zeros(2 * C * (L.BN_BLOCKS + 1), dtype=torch.float64, device="cuda")
    rm_d, rv_d = rm.clone(), rv.clone()
    a = L.BnStatsArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.z_cs, a.eps, a.momentum = dt, B, h, w, C, cs, 1e-5, 0.1
    a.z, a.mean_rstd, a.running_mean, a.running_var, a.scratch = za.data_ptr(), mr.data_ptr(), rm_d.data_ptr(), rv_d.data_ptr(), scratch.data_ptr()
    ctx().call("buctd_bn_stats", a, stream())
    y = torch.full_like(za, 3.0)
    b = L.BnApplyArgs()
    b.dtype, b.batch, b.h, b.w, b.c, b.z_cs, b.res_cs, b.out_cs, b.relu = dt, B, h, w, C, cs, cs, cs, 1
    b.z, b.mean_rstd, b.gamma, b.beta, b.residual, b.out = za.data_ptr(), mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(), resa.data_ptr(), y.data_ptr()
    ctx().call("buctd_bn_apply", b, stream())
    torch.cuda.synchronize()
    # reference: nn.BatchNorm2d in training mode
    zr = z.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    yr = F.relu(F.batch_norm(zr, rm_r, rv_r, gr, br, training=True, momentum=0.1, eps=1e-5) + rr)
    assert_close(mr[0], z.mean((0, 2, 3)), 1e-5, "batch mean")
    assert_close(mr[1], 1.0 / torch.sqrt(z.var((0, 2, 3), unbiased=False) + 1e-5), 1e-4, "rstd")
    assert_close(rm_d, rm_r, 1e-5, "running_mean")
    assert_close(rv_d, rv_r, 1e-5, "running_var")
    assert pads_are_zero(y) and (C == cs or y[..., C:].float().abs().max() == 0)
    assert_close(from_act16(y, C), yr.detach(), TOL[dt], "bn forward")
    # backward
    dy = q16(_rand(B, C, h, w, seed=16), dt)
    yr.backward(dy)
    dya = to_act16(dy, dt)
    dz = torch.full_like(za, 5.0)
    dres = torch.full_like(za, 5.0)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    c = L.BnBwdArgs()
    c.dtype, c.batch, c.h, c.w, c.c = dt, B, h, w, C
    c.dy_cs = c.y_cs = c.z_cs = c.dz_cs = c.dres_cs = cs
    c.relu, c.res_accumulate, c.inv_scale = 1, 0, 0.5
    c.dy, c.y, c.z, c.mean_rstd, c.gamma = dya.data_ptr(), y.data_ptr(), za.data_ptr(), mr.data_ptr(), gamma.data_ptr()
    c.dgamma, c.dbeta, c.scratch, c.dz, c.dres = dg.data_ptr(), db.data_ptr(), scratch.data_ptr(), dz.data_ptr(), dres.data_ptr()
    ctx().call("buctd_bn_bwd", c, stream())
    torch.cuda.synchronize()
    # the ReLU mask comes from the 16-bit y: where the fp32 reference is within rounding of 0 the masks may differ -> compare
    # against the reference evaluated with OUR mask
    mask = (from_act16(y, C) > 0).float()
    zr2 = z.clone().requires_grad_(True)
    g2, b2 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out2 = F.batch_norm(zr2, None, None, g2, b2, training=True, eps=1e-5)
    out2.backward(dy * mask)
    assert pads_are_zero(dz) and pads_are_zero(dres)
    assert_close(from_act16(dz, C), zr2.grad, TOL[dt], "bn dz")
    assert_close(from_act16(dres, C), dy * mask, 1e-6, "residual gradient")
    assert_close(dg, 0.5 * g2.grad, 2e-3, "dgamma")
    assert_close(db, 0.5 * b2.grad, 2e-3, "dbeta")
    c.res_accumulate = 1
    ctx().call("buctd_bn_bwd", c, stream())
    torch.cuda.synchronize()
    assert_close(from_act16(dres, C), 2 * dy * mask, TOL[dt], "residual gradient accumulated")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("C,rows", [(48, 1000), (6912, 96), (17, 333)])
def test_colsum(dt, C, rows):
    cs = (C + 7) // 8 * 8
    x = torch.zeros(rows, cs, dtype=TDT[dt], device="cuda")
    x[:, :C] = _rand(rows, C, seed=20).to(TDT[dt])
    out = torch.ones(C, device="cuda")
    scratch = torch.zeros(2048 * L.BN_BLOCKS, dtype=torch.float64, device="cuda")
    a = L.ColsumArgs()
    a.dtype, a.c, a.cs, a.accumulate, a.rows, a.inv_scale = dt, C, cs, 1, rows, 2.0
    a.dy, a.dbias, a.scratch = x.data_ptr(), out.data_ptr(), scratch.data_ptr()
    ctx().call("buctd_colsum", a, stream())
    torch.cuda.synchronize()
    assert_close(out, 1.0 + 2.0 * x[:, :C].float().sum(0), 1e-5, "colsum")


@pytest.mark.parametrize("dt", DTS)
def test_zero_stuff_and_stride2_gradients(dt):
    """data + weight gradient of a 3x3 stride-2 convolution through the zero-stuffed output gradient"""
    B, cin, cout, h, w = 2, 32, 48, 5, 4                      # output h x w, input 2h x 2w
    x = q16(_rand(B, cin, 2 * h, 2 * w, seed=21), dt)
    dy = q16(_rand(B, cout, h, w, seed=22), dt)
    wt = q16(_rand(cout, cin, 3, 3, seed=23) * 0.2, dt)
    dya = to_act16(dy, dt)
    up = torch.full((B, 2 * h + 1, 2 * w + 1, cout), 9.0, dtype=TDT[dt], device="cuda")
    a = L.ZeroStuffArgs()
    a.dtype, a.batch, a.h, a.w, a.cs = dt, B, h, w, cout
    a.in_, a.out = dya.data_ptr(), up.data_ptr()
    ctx().call("buctd_zero_stuff", a, stream())
    torch.cuda.synchronize()
    ref_up = torch.zeros(B, cout, 2 * h, 2 * w, device="cuda")
    ref_up[:, :, ::2, ::2] = dy
    assert pads_are_zero(up) and torch.equal(from_act16(up, cout), ref_up)
    # dx = stride-1 transposed convolution of the stuffed gradient = tap-GEMM with W'[8-t][ci][co]
    from gpu_utils import tapgemm
    wt_t = wt.permute(1, 0, 2, 3).flip(2, 3).contiguous()     # [cin][cout][3][3] flipped
    dx = torch.zeros(B, 2 * h + 1, 2 * w + 1, cin, dtype=TDT[dt], device="cuda")
    tapgemm(dt, up, wt_t, None, dx, H=2 * h, W=2 * w)
    xr = x.clone().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(xr, wr, stride=2, padding=1).backward(dy)
    assert_close(from_act16(dx, cin), xr.grad, TOL[dt], "stride-2 data gradient")
    xa = to_act16(x, dt)
    rows = B * (2 * h + 1) * (2 * w + 1)
    part = torch.zeros(1, 9, cout, cin, device="cuda")
    offs = [(ky - 1) * (2 * w + 1) + (kx - 1) for ky in range(3) for kx in range(3)]
    gemm16(dt, up, xa, cout, cin, rows, a_mn=1, b_mn=1, ntaps=9, tap_off=offs, d=part, ldd=cin, d_strides=(cout * cin, 0, 0), b_rows=rows)
    assert_close(part[0].permute(1, 2, 0).reshape(cout, cin, 3, 3), wr.grad, 3e-3, "stride-2 weight gradient")


@pytest.mark.parametrize("dt", DTS)
def test_fuse_backward(dt):
    B, C, h, w = 2, 48, 8, 8
    t0, t1, t2 = q16(_rand(B, C, h, w, seed=24), dt), q16(_rand(B, C, h // 2, w // 2, seed=25), dt), q16(_rand(B, C, h // 4, w // 4, seed=26), dt)
    r = [t.clone().requires_grad_(True) for t in (t0, t1, t2)]
    out = F.relu(r[0] + F.interpolate(r[1], scale_factor=2, mode="nearest") + F.interpolate(r[2], scale_factor=4, mode="nearest"))
    dout = q16(_rand(B, C, h, w, seed=27), dt)
    out.backward(dout)
    outa, douta = to_act16(q16(out.detach(), dt), dt), to_act16(dout, dt)
    mask = (q16(out.detach(), dt) > 0).float()
    d = [torch.full((B, (h >> s) + 1, (w >> s) + 1, C), 1.0, dtype=TDT[dt], device="cuda") for s in range(3)]
    pre = to_act16(torch.ones(B, C, h // 2, w // 2, device="cuda"), dt)
    d[1].copy_(pre)
    a = L.FuseBwdArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.cs, a.nterms, a.relu = dt, B, h, w, C, C, 3, 1
    a.dout, a.out = douta.data_ptr(), outa.data_ptr()
    for j in range(3):
        a.dterm[j], a.term_cs[j], a.term_shift[j], a.term_accumulate[j] = d[j].data_ptr(), C, j, int(j == 1)
    ctx().call("buctd_hr_fuse_bwd", a, stream())
    torch.cuda.synchronize()
    g = dout * mask
    assert_close(from_act16(d[0], C), g, 1e-6, "fuse bwd term 0")
    assert_close(from_act16(d[1], C), 1.0 + F.avg_pool2d(g, 2) * 4, TOL[dt], "fuse bwd term 1 (accumulated)")
    assert_close(from_act16(d[2], C), F.avg_pool2d(g, 4) * 16, TOL[dt], "fuse bwd term 2")
    assert all(pads_are_zero(t) for t in d)


@pytest.mark.parametrize("dt", DTS)
def test_relayout_add_rowdot(dt):
    B, C, h, w = 2, 48, 6, 5
    ps = 32
    x = q16(_rand(B, C, h, w, seed=28), dt)
    planar = torch.zeros(B, C, ps, dtype=TDT[dt], device="cuda"); planar[:, :, :h * w] = x.view(B, C, -1)
    act = torch.full((B, h + 1, w + 1, C), 4.0, dtype=TDT[dt], device="cuda")
    a = L.RelayoutArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.cs, a.p_stride, a.accumulate = dt, B, h, w, C, C, ps, 0
    a.in_, a.out = planar.data_ptr(), act.data_ptr()
    ctx().call("buctd_planar_to_act16", a, stream())
    torch.cuda.synchronize()
    assert pads_are_zero(act) and torch.equal(from_act16(act, C), x)
    back = torch.zeros_like(planar)
    a.in_, a.out = act.data_ptr(), back.data_ptr()
    ctx().call("buctd_act16_to_planar", a, stream())
    torch.cuda.synchronize()
    assert torch.equal(back, planar)
    y = q16(_rand(B, C, h, w, seed=29), dt)
    ya = to_act16(y, dt)
    s = L.Add16Args()
    s.dtype, s.cs_out, s.cs_a, s.c, s.rows, s.accumulate = dt, C, C, C, B * (h + 1) * (w + 1), 1
    s.a, s.out = ya.data_ptr(), act.data_ptr()
    ctx().call("buctd_add16", s, stream())
    torch.cuda.synchronize()
    assert_close(from_act16(act, C), x + y, TOL[dt], "add16")
    out = torch.zeros(B * (h + 1) * (w + 1), device="cuda")
    r = L.RowdotArgs()
    r.dtype, r.c, r.a_cs, r.b_cs, r.rows = dt, C, C, C, B * (h + 1) * (w + 1)
    xa = to_act16(x, dt)
    r.a, r.b, r.out = xa.data_ptr(), ya.data_ptr(), out.data_ptr()
    ctx().call("buctd_rowdot", r, stream())
    torch.cuda.synchronize()
    assert_close(out, (xa.float() * ya.float()).sum(-1).reshape(-1), 1e-5, "rowdot")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_chan_softmax_backward(dt, p):
    B, C, ld = 3, 48, 48
    S = _rand(B, C, C, seed=30)
    probs = torch.softmax(S * 0.2, -1).contiguous()
    da_d = _rand(B, C, C, seed=31)
    ds = torch.zeros(B, C, ld, dtype=TDT[dt], device="cuda")
    pd = torch.zeros(B, C, ld, dtype=TDT[dt], device="cuda")
    a = L.ChanSoftmaxBwdArgs()
    a.dtype, a.batch, a.c, a.ld, a.prob_stride, a.scale, a.drop_p, a.seed = dt, B, C, ld, C * C, 0.2, p, 77
    a.probs, a.da_d, a.ds, a.probs_d = probs.data_ptr(), da_d.data_ptr(), ds.data_ptr(), pd.data_ptr()
    ctx().call("buctd_chan_softmax_bwd", a, stream())
    torch.cuda.synchronize()
    keep = torch.from_numpy(np.stack([drop_mask(77, b, C, C, p) for b in range(B)])).cuda().float() if p > 0 else torch.ones(B, C, C, device="cuda")
    Sr = S.clone().requires_grad_(True)
    ad = torch.softmax(Sr * 0.2, -1) * keep / (1 - p)
    ad.backward(da_d)
    assert_close(pd, (probs * keep / (1 - p)), TOL[dt], "dropped probabilities")
    assert_close(ds, Sr.grad, TOL[dt], "chan soft-max backward")
    if p > 0:
        assert 0.85 < keep.mean().item() < 0.95


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("C,h,w", [(48, 12, 9), (96, 7, 5), (192, 5, 4)])
def test_position_attention_train_forward_backward(dt, p, C, h, w):
    """training-mode position attention: forward with the hash dropout + lse2, backward = rowdot + two gemm16 + the
    element-wise soft-max backward + three gemm16, against autograd of softmax(QK^T/sqrt(C)) -> dropout -> @ V"""
    B = 2
    rows = (h + 1) * (w + 1)
    ld = (rows + 7) // 8 * 8
    q, k, v = [q16(_rand(B, C, h, w, seed=s, scale=1.2), dt) for s in (32, 33, 34)]
    qa = to_act16(q, dt)
    kv = to_act16(torch.cat([k, v], 1), dt)
    o = torch.zeros(B, h + 1, w + 1, C, dtype=TDT[dt], device="cuda")
    lse = torch.zeros(B, rows, device="cuda")
    a = L.PosAttArgs()
    a.dtype, a.batch, a.h, a.w, a.c = dt, B, h, w, C
    a.q_cs, a.k_cs, a.v_cs, a.o_cs = C, 2 * C, 2 * C, C
    a.drop_p, a.drop_seed = p, 1234
    a.q, a.k, a.v, a.o, a.lse2 = qa.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * C, o.data_ptr(), lse.data_ptr()
    ctx().run(L.OP_POS_ATT, a, stream())
    torch.cuda.synchronize()
    # valid-pixel index p = y * w + x  <->  padded row r = y * (w + 1) + x
    ridx = (torch.arange(h)[:, None] * (w + 1) + torch.arange(w)[None, :]).reshape(-1).cuda()
    if p > 0:
        keep = torch.from_numpy(np.stack([drop_mask(1234, b, rows, rows, p) for b in range(B)])).cuda()
        keep = keep[:, ridx][:, :, ridx].float()
    else:
        keep = torch.ones(B, h * w, h * w, device="cuda")
    qr, kr, vr = [t.view(B, C, -1).transpose(1, 2).clone().requires_grad_(True) for t in (q, k, v)]
    att = torch.softmax(qr @ kr.transpose(1, 2) / math.sqrt(C), -1)
    outr = (att * keep / (1 - p)) @ vr
    ref = outr.detach().transpose(1, 2).reshape(B, C, h, w)
    assert_close(from_act16(o, C), ref, 2.5 * TOL[dt], "pos att train forward")
    lse_ref = torch.logsumexp(qr.detach() @ kr.detach().transpose(1, 2) / math.sqrt(C), -1) * math.log2(math.e)
    assert_close(lse.view(B, rows)[:, ridx], lse_ref, 2e-3, "lse2")
    # ---- backward ----
    do = q16(_rand(B, C, h, w, seed=35), dt)
    outr.backward(do.view(B, C, -1).transpose(1, 2))
    doa = to_act16(do, dt)
    delta = torch.zeros(B * rows, device="cuda")
    r = L.RowdotArgs()
    r.dtype, r.c, r.a_cs, r.b_cs, r.rows = dt, C, C, C, B * rows
    r.a, r.b, r.out = doa.data_ptr(), o.data_ptr(), delta.data_ptr()
    ctx().call("buctd_rowdot", r, stream())
    s = torch.zeros(B, rows, ld, device="cuda")
    dpd = torch.zeros(B, rows, ld, device="cuda")
    sc = math.log2(math.e) / math.sqrt(C)
    kmat = kv                                      # K = columns [0, C), V = [C, 2C) of the kv tensor
    vview = kv.view(B * rows, 2 * C)[:, C:]
    gemm16(dt, qa, kmat, rows, rows, C, a_mn=0, b_mn=0, nbatch=B, a_bs=rows * C, b_bs=rows * 2 * C, d=s, ldd=ld, d_strides=(0, 0, rows * ld),
           alpha=sc, lda=C, ldb=2 * C)
    g = L.Gemm16Args()
    gemm16(dt, doa, vview, rows, rows, C, a_mn=0, b_mn=0, nbatch=B, a_bs=rows * C, b_bs=rows * 2 * C, d=dpd, ldd=ld, d_strides=(0, 0, rows * ld),
           lda=C, ldb=2 * C)
    pd = torch.zeros(B, rows, ld, dtype=TDT[dt], device="cuda")
    ds = torch.zeros(B, rows, ld, dtype=TDT[dt], device="cuda")
    e = L.PosSoftmaxBwdArgs()
    e.dtype, e.batch, e.h, e.w, e.ld, e.inv_sqrt_c, e.drop_p, e.seed = dt, B, h, w, ld, 1.0 / math.sqrt(C), p, 1234
    e.s, e.dpd, e.lse2, e.delta, e.pd, e.ds = s.data_ptr(), dpd.data_ptr(), lse.data_ptr(), delta.data_ptr(), pd.data_ptr(), ds.data_ptr()
    ctx().call("buctd_pos_softmax_bwd", e, stream())
    # the fused score stage (both products in TMEM, soft-max backward in the epilogue) must give the same pd / ds
    pd2 = torch.full_like(pd, 3.0)
    ds2 = torch.full_like(ds, 3.0)
    f = L.PosBwdScoresArgs()
    f.dtype, f.batch, f.h, f.w, f.c, f.q_cs, f.k_cs, f.v_cs, f.do_cs, f.ld = dt, B, h, w, C, C, 2 * C, 2 * C, C, ld
    f.drop_p, f.seed = p, 1234
    f.q, f.k, f.v, f.d_o = qa.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * C, doa.data_ptr()
    f.lse2, f.delta, f.pd, f.ds = lse.data_ptr(), delta.data_ptr(), pd2.data_ptr(), ds2.data_ptr()
    ctx().call("buctd_pos_att_bwd_scores", f, stream())
    torch.cuda.synchronize()
    assert_close(pd2, pd, 2e-3, "fused pd")
    assert_close(ds2, ds, 2e-3, "fused ds")
    assert (pd2 == 0)[:, :, rows:].all() and (ds2 == 0)[:, :, rows:].all()
    dkv = torch.zeros(B, h + 1, w + 1, 2 * C, dtype=TDT[dt], device="cuda")
    dq = torch.zeros(B, h + 1, w + 1, C, dtype=TDT[dt], device="cuda")
    dvview = dkv.view(B * rows, 2 * C)[:, C:]
    gemm16(dt, pd, doa, rows, C, rows, a_mn=1, b_mn=1, nbatch=B, a_bs=rows * ld, b_bs=rows * C, d=dvview, out_f32=False, ldd=2 * C,
           d_strides=(0, 0, rows * 2 * C), lda=ld, ldb=C)
    gemm16(dt, ds, qa, rows, C, rows, a_mn=1, b_mn=1, nbatch=B, a_bs=rows * ld, b_bs=rows * C, d=dkv, out_f32=False, ldd=2 * C,
           d_strides=(0, 0, rows * 2 * C), lda=ld, ldb=C)
    gemm16(dt, ds, kmat, rows, C, rows, a_mn=0, b_mn=1, nbatch=B, a_bs=rows * ld, b_bs=rows * 2 * C, d=dq, out_f32=False, ldd=C,
           d_strides=(0, 0, rows * C), lda=ld, ldb=2 * C)
    torch.cuda.synchronize()
    tol = 4 * TOL[dt]
    assert_close(from_act16(dq, C), qr.grad.transpose(1, 2).reshape(B, C, h, w), tol, "dQ")
    assert_close(from_act16(dkv[..., :C], C), kr.grad.transpose(1, 2).reshape(B, C, h, w), tol, "dK")
    assert_close(from_act16(dkv[..., C:], C), vr.grad.transpose(1, 2).reshape(B, C, h, w), tol, "dV")
    assert pads_are_zero(dq) and pads_are_zero(dkv)


@pytest.mark.parametrize("dt", DTS)
def test_channel_attention_dropout_forward(dt):
    B, C, h, w, p = 2, 48, 9, 7, 0.1
    y = q16(_rand(B, C, h, w, seed=36), dt)
    yc = q16(_rand(B, C, h, w, seed=37), dt)
    out = torch.zeros(B, C, h * w, dtype=TDT[dt], device="cuda")
    nch = L.chan_att_scratch_floats(B, h, w, C) // (B * C * C)
    scratch = torch.zeros(B, nch, C, C, dtype=torch.float32, device="cuda")
    ya, yca = to_act16(y, dt), to_act16(yc, dt)
    a = L.ChanAttArgs()
    a.dtype, a.batch, a.h, a.w, a.c, a.y_cs, a.yc_cs = dt, B, h, w, C, C, C
    a.drop_p, a.drop_seed = p, 99
    a.y, a.yc, a.scratch, a.out = ya.data_ptr(), yca.data_ptr(), scratch.data_ptr(), out.data_ptr()
    ctx().run(L.OP_CHAN_ATT, a, stream())
    torch.cuda.synchronize()
    yf, ycf = y.view(B, C, -1), yc.view(B, C, -1)
    att = torch.softmax(torch.matmul(ycf, yf.transpose(1, 2)) / math.sqrt(h * w), -1)
    keep = torch.from_numpy(np.stack([drop_mask(99, b, C, C, p) for b in range(B)])).cuda().float()
    assert_close(scratch[:, 0], att, 1e-4, "chan att probs (undropped)")
    assert_close(out, torch.matmul(att * keep / (1 - p), yf), TOL[dt], "chan att out with dropout")


@pytest.mark.parametrize("dt", DTS)
def test_stem_train_and_weight_gradient(dt):
    B, H, W = 2, 16, 12
    x = _rand(B, 5, H, W, seed=38)
    wt = _rand(64, 3, 3, 3, seed=39) * 0.3
    out = torch.zeros(B, H // 2 + 1, W // 2 + 1, 64, dtype=TDT[dt], device="cuda")
    w27, b0 = wt.reshape(64, 27).contiguous(), torch.zeros(64, device="cuda")
    a = L.StemArgs()
    a.dtype, a.batch, a.x_channels, a.h, a.w, a.out_cs, a.x_half, a.no_relu = dt, B, 5, H, W, 64, 0, 1
    a.x, a.weight, a.bias, a.out = x.data_ptr(), w27.data_ptr(), b0.data_ptr(), out.data_ptr()
    ctx().run(L.OP_STEM, a, stream())
    torch.cuda.synchronize()
    wr = wt.clone().requires_grad_(True)
    ref = F.conv2d(x[:, :3], wr, stride=2, padding=1)
    assert_close(from_act16(out, 64), ref.detach(), TOL[dt], "stem raw output (no ReLU)")
    assert (from_act16(out, 64) < 0).any()
    dz = q16(_rand(B, 64, H // 2, W // 2, seed=40), dt)
    ref.backward(dz)
    dza = to_act16(dz, dt)
    part = torch.zeros(L.BN_BLOCKS, 64, 27, device="cuda")
    s = L.StemWgradArgs()
    s.dtype, s.batch, s.x_channels, s.h, s.w, s.dz_cs, s.x_half = dt, B, 5, H, W, 64, 0
    s.x, s.dz, s.partial = x.data_ptr(), dza.data_ptr(), part.data_ptr()
    ctx().call("buctd_stem_wgrad", s, stream())
    grad = torch.zeros(64, 3, 3, 3, device="cuda")
    r = L.WgradReduceArgs()
    r.n, r.k, r.ntaps, r.nsplit, r.accumulate, r.inv_scale = 64, 27, 1, L.BN_BLOCKS, 0, 1.0
    r.partial, r.grad = part.data_ptr(), grad.data_ptr()
    ctx().call("buctd_wgrad_reduce", r, stream())
    torch.cuda.synchronize()
    assert_close(grad, wr.grad, 1e-4, "stem weight gradient")


@pytest.mark.parametrize("dt", DTS)
def test_repack_weights(dt):
    convs = [(48, 32, 9), (17, 48, 1), (96, 96, 9)]
    params, table, dst_off = [], [], 0
    src_off = 0
    exp = []
    for cout, cin, taps in convs:
        w = _rand(cout, cin, taps, seed=41 + cout)
        params.append(w.reshape(-1))
        for tr in (0, 1):
            rows_alloc = (cout if not tr else cin) + 15 & ~15
            cs = ((cin if not tr else cout) + 7) // 8 * 8
            table.append([src_off, dst_off, cout, cin, taps, rows_alloc, cs, tr])
            e = torch.zeros(taps, rows_alloc, cs, device="cuda")
            if not tr:
                e[:, :cout, :cin] = w.permute(2, 0, 1)
            else:
                e[:, :cin, :cout] = w.permute(2, 1, 0).flip(0)
            exp.append((dst_off, e))
            dst_off += taps * rows_alloc * cs
        src_off += w.numel()
    flat = torch.cat(params)
    w16 = torch.full((dst_off,), 3.0, dtype=TDT[dt], device="cuda")
    tab = torch.tensor(table, dtype=torch.int64, device="cuda")
    units = [[e, i, j, 0] for e, t in enumerate(table) for i in range((t[2] + 31) // 32) for j in range((t[3] + 31) // 32)]
    un = torch.tensor(units, dtype=torch.int32, device="cuda")
    w16.zero_()                                      # padding rows / columns are not written by the kernel
    a = L.RepackArgs()
    a.dtype, a.n_entries, a.n_units, a.param, a.w16, a.table, a.units = dt, len(table), len(units), flat.data_ptr(), w16.data_ptr(), tab.data_ptr(), un.data_ptr()
    ctx().call("buctd_repack_weights", a, stream())
    torch.cuda.synchronize()
    for off, e in exp:
        assert torch.equal(w16[off:off + e.numel()].view_as(e), e.to(TDT[dt]))
