"""-m gpu: decode + rasteriser kernels through the reference-facing Python API (which calls the C
ABI) against the oracle and the committed reference outputs (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from buctd_b200 import condition, inference
from buctd_b200.config import make_cfg
from oracle import condition_oracle as C
from oracle import decode_oracle as D
from oracle import make_golden as G

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", list(G.CASES) + ["decode_edge_cases"])
def test_decode_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    cfg = make_cfg("plain", 32, (72, 96))
    coords, maxvals = inference.get_max_preds(g["heatmaps"])
    assert np.array_equal(coords, g["coords"])              # integer index work: bit exact
    assert np.array_equal(maxvals, g["maxvals"])
    for pp in (False, True):
        key = "preds_pp%d" % pp if name == "decode_edge_cases" else ("preds" if pp else None)
        if key is None:
            continue
        cfg.TEST.POST_PROCESS = pp
        preds, mv = inference.get_final_preds(cfg, g["heatmaps"], g["center"], g["scale"])
        assert np.array_equal(mv, g["maxvals"])
        assert np.abs(preds - g[key]).max() <= 1e-4         # float64 affine stored as float32


def test_decode_device_indices_and_offsets_vs_oracle():
    g = np.random.default_rng(5)
    B, K, H, W = 64, 17, 96, 72                              # BASELINE.json full size
    hm = g.standard_normal((B, K, H, W)).astype(np.float32)
    hm[3, 2] = hm[3, 2].round(1)                             # many exact ties
    hm[4, :] = -np.abs(hm[4, :])                             # all negative
    center = g.uniform(100, 200, (B, 2)).astype(np.float32)
    scale = g.uniform(1, 2, (B, 2)).astype(np.float32)
    t = torch.from_numpy(hm).cuda()
    preds, maxvals, coords, idx = inference.decode_device(t, torch.from_numpy(center).cuda(), torch.from_numpy(scale).cuda(),
                                                          post_process=True, want_coords=True)
    oc, om = D.get_max_preds(hm)
    op, _ = D.get_final_preds(True, hm, center, scale)
    assert np.array_equal(idx.cpu().numpy(), hm.reshape(B, K, -1).argmax(-1).astype(np.int32))
    assert np.array_equal(coords.cpu().numpy(), oc)
    assert np.array_equal(maxvals.cpu().numpy(), om)
    assert np.abs(preds.cpu().numpy() - op).max() <= 1e-4
    # heat-map-space result (no center/scale): quarter offsets are exact in float32
    p2, _ = inference.decode_device(t, None, None, post_process=True)
    ident_c = np.tile(np.array([[W / 2.0, H / 2.0]], np.float32), (B, 1))
    ident_s = np.tile(np.array([[W / 200.0, H / 200.0]], np.float32), (B, 1))
    op2, _ = D.get_final_preds(True, hm, ident_c, ident_s)
    assert np.abs(p2.cpu().numpy() - op2).max() <= 1e-4


def test_decode_empty_batch_and_asserts():
    cfg = make_cfg("plain", 32, (72, 96))
    p, m = inference.get_final_preds(cfg, np.zeros((0, 17, 8, 6), np.float32), np.zeros((0, 2)), np.zeros((0, 2)))
    assert p.shape == (0, 17, 2) and m.shape == (0, 17, 1)
    with pytest.raises(AssertionError):
        inference.get_max_preds(np.zeros((17, 8, 6), np.float32))
    with pytest.raises(AssertionError):
        inference.get_max_preds([[1.0]])


@pytest.mark.parametrize("mode", ["stacked", "colored", "plain"])
def test_rasteriser_vs_oracle(mode):
    g = np.random.default_rng(9)
    B, K, H, W = 3, 17, 384, 288
    kp = np.stack([g.uniform(-10, W + 10, (B, K)), g.uniform(-10, H + 10, (B, K)), g.uniform(0, 1, (B, K))], -1).astype(np.float32)
    kp[0, 0] = (1.2, 1.7, 1)          # -> pixel (0,0): reflect-101 corner
    kp[0, 1] = (W - 0.5, H - 0.5, 1)  # -> last valid pixel (W-2,H-2)
    kp[0, 2] = (0.5, 30, 1)           # x<1: skipped
    kp[0, 3] = kp[0, 4]               # duplicate position: later keypoint overwrites
    kp[1, :] = 0                      # no keypoint at all: all-zero image, no division by zero
    out = condition.rasterize_condition(torch.from_numpy(kp).cuda(), H, W, mode).cpu().numpy()
    for b in range(B):
        if mode == "stacked":
            ref = C.stacked_condition(kp[b], H, W)
        elif mode == "colored":
            ref = C.colored_condition(kp[b], H, W, C.default_colors(K))
        else:
            ref = C.plain_condition(kp[b], H, W)
        ref = ref.astype(np.float32)
        if mode == "plain":
            # .astype(int) truncation: a value within 1e-9 of an integer may land on either side
            bad = np.abs(out[b] - ref) > 0
            assert bad.mean() < 1e-4 and np.abs(out[b] - ref).max() <= 1.0
        else:
            assert np.abs(out[b] - ref).max() <= 2e-4, (mode, b, np.abs(out[b] - ref).max())
    assert out[1].max() == 0


@pytest.mark.parametrize("name", ["s128x96", "s384x288", "s40x30", "all_skipped", "single"])
def test_rasteriser_vs_reference_source_golden(golden_dir, name):
    """The CUDA rasteriser against the outputs of the reference's own rasteriser source (tests/golden/condition.npz,
    oracle/make_golden_condition.py): all three modes, borders / skips / duplicates."""
    import os
    g = np.load(os.path.join(golden_dir, "condition.npz"))
    kp = torch.from_numpy(g[name + ":kpts"].astype(np.float32)[None]).cuda()
    h, w = [int(v) for v in g[name + ":size"]]
    K = kp.shape[1]
    cols = [list(g["colors"][i % len(g["colors"])]) for i in range(K)]
    st = condition.rasterize_condition(kp, h, w, "stacked").cpu().numpy()[0]
    assert np.abs(st - g[name + ":stacked"]).max() <= 2e-4
    co = condition.rasterize_condition(kp, h, w, "colored", colors=cols).cpu().numpy()[0]
    assert np.abs(co - g[name + ":colored"]).max() <= 2e-4
    pl = condition.rasterize_condition(kp, h, w, "plain").cpu().numpy()[0]
    d = np.abs(pl - g[name + ":plain"].astype(np.float32))
    # .astype(int) truncation of a float64 blur: the float32 device value may land on the other side of an integer
    assert d.max() <= 1.0 and (d > 0).mean() < 1e-4


def test_rasteriser_writes_into_model_input_channels():
    B, K, H, W = 2, 17, 128, 96
    kp = torch.tensor(np.random.default_rng(1).uniform(5, 90, (B, K, 2)), dtype=torch.float32).cuda()
    x = torch.full((B, 3 + K, H, W), -1.0, device="cuda")
    condition.rasterize_condition(kp, H, W, "stacked", out=x, c_off=3)
    assert (x[:, :3] == -1).all() and x[:, 3:].max() > 254.9 and x[:, 3:].min() >= 0


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("H,W", [(384, 288), (128, 96), (256, 192)])
def test_cond_pyramid_from_proposals(dt, H, W):
    """buctd_cond_from_kpts == rasterise (stacked) -> antialiased resize to the three CoAM scales, without the planes:
    against the device composition (same real values, independent float paths: <= 1 ulp of the 16-bit type at 255) and
    against the oracle composition (condition_oracle -> hrnet_oracle.resize_aa) within 1e-3 of 255."""
    from buctd_b200 import _lib as L
    from oracle import hrnet_oracle as O
    tdt = torch.float16 if dt == 0 else torch.bfloat16
    g = np.random.default_rng(21)
    B, K = 3, 17
    kp = np.stack([g.uniform(-6, W + 6, (B, K)), g.uniform(-6, H + 6, (B, K)), g.uniform(0, 1, (B, K))], -1).astype(np.float32)
    kp[0, 0] = (1.2, 1.7, 1); kp[0, 1] = (W - 0.5, H - 0.5, 1); kp[0, 2] = (0.5, 30, 1); kp[0, 3] = (3.3, H - 2.2, 1)
    kp[0, 4] = (W - 3.0, 5.1, 1); kp[1, :] = 0
    kpd = torch.from_numpy(kp).cuda()
    ctx = L.Context.get(0)
    st = torch.cuda.current_stream().cuda_stream
    outs, refs = [], []
    a = L.CondKptsArgs()
    a.dtype, a.batch, a.k, a.kstride, a.h, a.w = dt, B, K, 3, H, W
    r3 = L.Resize3Args()
    x = condition.rasterize_condition(kpd, H, W, "stacked")
    r3.dtype, r3.batch, r3.x_channels, r3.c_off, r3.cc, r3.h, r3.w = dt, B, K, 0, K, H, W
    for i in range(3):
        oh, ow = H >> (2 + i), W >> (2 + i)
        outs.append(torch.full((B, oh + 1, ow + 1, 32), 7.0, dtype=tdt, device="cuda"))
        refs.append(torch.zeros(B, oh + 1, ow + 1, 32, dtype=tdt, device="cuda"))
        a.out_cs[i], a.out[i] = 32, outs[i].data_ptr()
        r3.out_cs[i], r3.out[i] = 32, refs[i].data_ptr()
    a.kpts, r3.x = kpd.data_ptr(), x.data_ptr()
    ctx.call("buctd_cond_from_kpts", a, st)
    ctx.call("buctd_cond_resize3", r3, st)
    torch.cuda.synchronize()
    ulp = 0.125 if dt == 0 else 1.0          # spacing of the 16-bit type in [128, 256)
    xo = torch.from_numpy(np.stack([C.stacked_condition(kp[b], H, W) for b in range(B)]).astype(np.float32))
    for i in range(3):
        o, r = outs[i].float(), refs[i].float()
        assert o[:, -1].abs().max() == 0 and o[:, :, -1].abs().max() == 0 and o[..., K:].abs().max() == 0     # pads, spare channels
        assert (o - r).abs().max().item() <= ulp, (i, (o - r).abs().max().item())
        assert ((o - r).abs() > 0).float().mean().item() < 0.02
        oh, ow = H >> (2 + i), W >> (2 + i)
        ora = O.resize_aa(xo, oh, ow).permute(0, 2, 3, 1)
        tol = 1e-3 * 255 + (ulp / 2)
        assert (o[:, :oh, :ow, :K].cpu() - ora).abs().max().item() <= tol
    assert outs[0][1].abs().max() == 0                                       # no proposals: all-zero condition
