"""Helpers for the -m gpu tests: build act16 tensors with torch (test infrastructure only) and call
the C ABI directly."""
import numpy as np
import torch

from buctd_b200 import _lib as L

TDT = {L.F16: torch.float16, L.BF16: torch.bfloat16}


def rup(v, m):
    return (v + m - 1) // m * m


def stream():
    return torch.cuda.current_stream().cuda_stream


def ctx():
    return L.Context.get(torch.cuda.current_device())


def to_act16(x, dt, cs=None):
    """x [B,C,H,W] float32 (cuda) -> act16 tensor [B,H+1,W+1,cs] with zero pads."""
    B, C, H, W = x.shape
    cs = cs or rup(C, 8)
    a = torch.zeros(B, H + 1, W + 1, cs, dtype=TDT[dt], device=x.device)
    a[:, :H, :W, :C] = x.permute(0, 2, 3, 1).to(TDT[dt])
    return a


def from_act16(a, C, H=None, W=None):
    B, Hp, Wp, cs = a.shape
    return a[:, :Hp - 1, :Wp - 1, :C].permute(0, 3, 1, 2).float().contiguous()


def pads_are_zero(a):
    return bool((a[:, -1].float().abs().max() == 0) and (a[:, :, -1].float().abs().max() == 0))


def pack_w(w4, dt):
    """[cout,cin,kh,kw] float -> [taps][n_alloc][k_cs] 16-bit (cuda)"""
    cout, cin, kh, kw = w4.shape
    n_alloc, k_cs = rup(cout, 16), rup(cin, 8)
    p = torch.zeros(kh * kw, n_alloc, k_cs, dtype=torch.float32, device=w4.device)
    p[:, :cout, :cin] = w4.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    return p.to(TDT[dt]).contiguous(), n_alloc, k_cs


def tapgemm(dt, a16, w4, bias, out, *, H, W, stride=1, relu=False, residual=None, out_mode=L.OUT_ACT16, k=None,
            m_rows=None, in_hw=None, batch=None):
    wp, n_alloc, k_cs = pack_w(w4, dt)
    g = L.TapGemmArgs()
    g.dtype, g.mode, g.ntaps, g.stride = dt, (L.TG_PATCH if stride != 1 else L.TG_FLAT), w4.shape[2] * w4.shape[3], stride
    g.batch = batch if batch is not None else a16.shape[0]
    g.out_h, g.out_w = H, W
    g.in_h, g.in_w = in_hw if in_hw else (H, W)
    g.m_rows = m_rows if m_rows is not None else g.batch * (H + 1) * (W + 1)
    g.k = k if k is not None else w4.shape[1]
    g.a_cs = a16.shape[-1]
    g.n, g.n_alloc, g.w_cs = w4.shape[0], n_alloc, k_cs
    g.out_cs = out.shape[-1] if out_mode != L.OUT_NCHW32 else 0
    g.res_cs = residual.shape[-1] if residual is not None else 0
    g.relu, g.out_mode = int(relu), out_mode
    b32 = bias.float().contiguous() if bias is not None else None
    g.a, g.w, g.bias = a16.data_ptr(), wp.data_ptr(), (b32.data_ptr() if b32 is not None else None)
    g.residual = residual.data_ptr() if residual is not None else None
    g.out = out.data_ptr()
    ctx().run(L.OP_TAPGEMM, g, stream())
    torch.cuda.synchronize()
    return out


def q16(x, dt):
    """round a float tensor through the 16-bit type (what the kernels see)"""
    return x.to(TDT[dt]).float()


def assert_close(got, ref, rel=4e-3, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, "%s: max-abs err %.3e > %.1e * %.3e" % (what, err, rel, scale)
    return err
