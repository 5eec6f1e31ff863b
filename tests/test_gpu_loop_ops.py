"""-m gpu: the kernels either side of the forward (flip test, JointsMSELoss, Adam, iterative refinement) through the
C ABI / host mirrors, against oracle/loop_oracle.py and the reference-generated tests/golden/loop_ops.npz."""
import os

import numpy as np
import pytest
import torch

from buctd_b200 import condition, function, inference, loss, optim, transforms
from oracle import condition_oracle as CO
from oracle import decode_oracle as D
from oracle import loop_oracle as LO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "loop_ops.npz"))


def test_flip_back_golden(g):
    pairs = g["flip_pairs"].tolist()
    assert np.array_equal(transforms.flip_back(g["hm_flipped"].copy(), pairs), g["flip_back"])          # bit exact: pure data movement


def test_flip_input_golden(g):
    out = transforms.flip_input_device(torch.from_numpy(g["x"]).cuda(), g["flip_pairs"].tolist(), stacked=True)
    assert np.array_equal(out.cpu().numpy(), g["x_flipped"])


@pytest.mark.parametrize("shift", [0, 1])
def test_flip_merge_golden(g, shift):
    out = transforms.flip_merge_device(torch.from_numpy(g["hm"]).cuda(), torch.from_numpy(g["hm_flipped"]).cuda(),
                                       g["flip_pairs"].tolist(), bool(shift))
    assert np.array_equal(out.cpu().numpy(), g["merge_shift%d" % shift])                             # (a + b) * 0.5 in fp32: exact


@pytest.mark.parametrize("shape", [(1, 17, 96, 72), (5, 14, 7, 33), (64, 17, 96, 72)])
def test_flip_merge_shapes(shape):
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32)
    pairs = [[0, 1], [2, 5], [3, 4]] if shape[1] == 14 else function.COCO_FLIP_PAIRS
    out = transforms.flip_merge_device(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), pairs, True)
    assert np.array_equal(out.cpu().numpy(), LO.flip_test_merge(a, b, pairs, True))


@pytest.mark.parametrize("utw", [0, 1])
def test_joints_mse_golden(g, utw):
    o = torch.from_numpy(g["loss_output"]).cuda().requires_grad_(True)
    crit = loss.JointsMSELoss(bool(utw))
    val = crit(o, torch.from_numpy(g["loss_target"]).cuda(), torch.from_numpy(g["loss_weight"]).cuda())
    val.backward()
    ref = float(g["loss_utw%d" % utw])
    assert abs(val.item() - ref) <= 1e-6 * max(1.0, abs(ref))          # fp32 value, fp64 accumulation
    assert np.abs(o.grad.cpu().numpy() - g["grad_utw%d" % utw]).max() <= 1e-9


def test_joints_mse_full_size_matches_oracle():
    rng = np.random.default_rng(9)
    o = rng.standard_normal((32, 17, 96, 72)).astype(np.float32)
    t = rng.uniform(0, 1, o.shape).astype(np.float32)
    w = (rng.uniform(size=(32, 17, 1)) < 0.8).astype(np.float32)
    val = loss.JointsMSELoss(True)(torch.from_numpy(o).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(w).cuda()).item()
    ref = D.joints_mse_loss(o, t, w, True)
    assert abs(val - ref) <= 1e-6 * abs(ref)


def test_adam_golden(g):
    p = torch.nn.Parameter(torch.from_numpy(g["adam_p0"].copy()).cuda())
    opt = optim.Adam([p], lr=1e-3)
    for i in range(3):
        p.grad.copy_(torch.from_numpy(g["adam_g%d" % i]).cuda())
        opt.step()
        assert np.abs(p.detach().cpu().numpy() - g["adam_p%d" % (i + 1)]).max() <= 2e-7
    n = g["adam_p0"].size
    assert np.abs(opt.exp_avg[:n].cpu().numpy() - g["adam_m3"]).max() <= 1e-7
    assert np.abs(opt.exp_avg_sq[:n].cpu().numpy() - g["adam_v3"]).max() <= 1e-6


def test_adam_many_tensors_matches_torch():
    """flat-buffer re-homing: ragged tensor sizes, 10 steps, against torch.optim.Adam on the same device"""
    torch.manual_seed(0)
    shapes = [(48, 48, 3, 3), (48,), (17, 48, 1, 1), (5,), (6912, 13), (1,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    a, b = optim.Adam(ours, lr=2e-3), torch.optim.Adam(theirs, lr=2e-3)
    for step in range(10):
        for p, q in zip(ours, theirs):
            gr = torch.randn_like(q) * (0.1 + step)
            p.grad.copy_(gr)
            q.grad = gr.clone()
        a.step(); b.step()
    for p, q in zip(ours, theirs):
        assert (p.detach() - q.detach()).abs().max().item() <= 1e-6


def test_iterative_refine_feeds_back_decoded_keypoints():
    """configs[3]: the condition of pass n+1 is the rasterised decode of pass n (checked with the oracle rasteriser and
    the oracle decode applied to the GPU's own pass-n heat-maps: identical inputs on both sides)."""
    from buctd_b200 import pose_hrnet_coam
    from oracle import make_golden
    name = "coam_w32_128x96_stacked"
    cfg, _ = make_golden.case_cfg(name)
    cfg.TEST.POST_PROCESS = True
    net = pose_hrnet_coam.get_pose_net(cfg, is_train=False)
    net.load_state_dict(make_golden.case_state_dict(name), strict=True)
    net = net.cuda().eval()
    x = torch.from_numpy(make_golden.case_inputs(name)).cuda()
    outs = function.iterative_refine(cfg, net, x, passes=3, return_all=True)
    assert len(outs) == 3
    # replay pass 1 -> pass 2 by hand
    with torch.no_grad():
        hm1 = net(x)
    c1, _ = D.get_final_preds(True, hm1.cpu().numpy(), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32))
    coords1, _ = inference.decode_device(hm1, None, None, post_process=True)
    assert np.array_equal(outs[0][0].cpu().numpy(), coords1.cpu().numpy())
    kp = LO.refine_condition(coords1.cpu().numpy())
    cond = np.stack([CO.stacked_condition(np.concatenate([kp[b], np.ones((17, 1), np.float32)], 1), 128, 96) for b in range(2)])
    x2 = x.clone()
    x2[:, 3:] = torch.from_numpy(cond.astype(np.float32)).cuda()
    with torch.no_grad():
        hm2 = net(x2)
    coords2, _ = inference.decode_device(hm2, None, None, post_process=True)
    assert np.abs(outs[1][0].cpu().numpy() - coords2.cpu().numpy()).max() <= 0.25 + 1e-6      # raster tolerance may flip a quarter offset


def test_forward_flip_test_matches_manual_composition():
    from buctd_b200 import pose_hrnet_coam
    from oracle import make_golden
    name = "coam_w32_128x96_stacked"
    cfg, _ = make_golden.case_cfg(name)
    cfg.TEST.FLIP_TEST, cfg.TEST.SHIFT_HEATMAP = True, True
    net = pose_hrnet_coam.get_pose_net(cfg, is_train=False)
    net.load_state_dict(make_golden.case_state_dict(name), strict=True)
    net = net.cuda().eval()
    xn = make_golden.case_inputs(name)
    x = torch.from_numpy(xn).cuda()
    out = function.forward_flip_test(cfg, net, x).cpu().numpy()
    with torch.no_grad():
        o = net(x).cpu().numpy()
        of = net(torch.from_numpy(LO.flip_test_input(xn, function.COCO_FLIP_PAIRS, stacked=True)).cuda()).cpu().numpy()
    assert np.array_equal(out, LO.flip_test_merge(o, of, function.COCO_FLIP_PAIRS, True))


@pytest.fixture(scope="module")
def gt(golden_dir):
    return np.load(os.path.join(golden_dir, "train_ops.npz"))


@pytest.mark.parametrize("name", ["w48_384x288", "w32_256x192_s2", "odd_stride"])
def test_generate_target_bit_exact(gt, name):
    """device targets == the reference's JointsDataset.generate_target (golden written from the reference's own source)"""
    from buctd_b200 import evaluate as E
    img_w, img_h, hm_w, hm_h, sigma = gt["tg_%s_meta" % name].tolist()
    jw = gt["tg_%s_jw" % name] if ("tg_%s_jw" % name) in gt.files else None
    t, w = E.generate_target(gt["tg_%s_joints" % name], gt["tg_%s_vis" % name], (img_w, img_h), (hm_w, hm_h), sigma, jw)
    assert t.is_cuda and t.dtype == torch.float32
    assert np.array_equal(t.cpu().numpy(), gt["tg_%s_target" % name])
    assert np.array_equal(w.cpu().numpy(), gt["tg_%s_weight" % name])


def test_generate_target_random_vs_oracle():
    from buctd_b200 import evaluate as E
    from oracle import loop_oracle as LO
    g = np.random.default_rng(77)
    B, K = 32, 17
    j = np.zeros((B, K, 3))
    j[:, :, 0], j[:, :, 1] = g.uniform(-60, 350, (B, K)), g.uniform(-60, 450, (B, K))
    v = np.repeat((g.uniform(size=(B, K, 1)) < 0.8).astype(np.float64), 3, axis=2)
    t, w = E.generate_target(j, v, (288, 384), (72, 96), 3)
    for b in range(B):
        to, wo = LO.generate_target(j[b], v[b], (288, 384), (72, 96), 3)
        assert np.array_equal(t[b].cpu().numpy(), to) and np.array_equal(w[b].cpu().numpy(), wo)


def test_accuracy_bit_exact(gt):
    from buctd_b200 import evaluate as E
    acc, avg, cnt, pred = E.accuracy(gt["acc_output"], gt["acc_target"])
    assert np.array_equal(acc, gt["acc_acc"]) and avg == float(gt["acc_avg"]) and cnt == int(gt["acc_cnt"])
    assert np.array_equal(pred, gt["acc_pred"])
    acc0, avg0, cnt0, _ = E.accuracy(gt["acc_output"][:2], np.zeros_like(gt["acc_target"][:2]))
    assert np.array_equal(acc0, gt["acc0_acc"]) and avg0 == 0 and cnt0 == 0
