"""not gpu: pins oracle/loop_oracle.py (flip test, JointsMSELoss gradient, Adam) against tests/golden/loop_ops.npz,
written by oracle/make_golden_loop.py from the reference's own flip_back / flip_hm / JointsMSELoss (autograd) and
torch.optim.Adam; where /root/reference is mounted the reference functions are also re-run live."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from oracle import loop_oracle as LO
from oracle import refharness as R


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "loop_ops.npz"))


def test_flip_back_and_input_flip(g):
    pairs = g["flip_pairs"].tolist()
    assert np.array_equal(LO.flip_back(g["hm_flipped"], pairs), g["flip_back"])
    assert np.array_equal(LO.flip_test_input(g["x"], pairs, stacked=True), g["x_flipped"])
    perm = LO.pairs_to_perm(17, pairs)
    assert np.array_equal(g["hm_flipped"][:, perm][:, :, :, ::-1], g["flip_back"])       # perm form == sequential swaps


@pytest.mark.parametrize("shift", [0, 1])
def test_flip_merge(g, shift):
    out = LO.flip_test_merge(g["hm"], g["hm_flipped"], g["flip_pairs"].tolist(), bool(shift))
    assert np.array_equal(out, g["merge_shift%d" % shift])


@pytest.mark.parametrize("utw", [0, 1])
def test_joints_mse_value_and_grad(g, utw):
    o, t, w = g["loss_output"], g["loss_target"], g["loss_weight"]
    assert abs(D.joints_mse_loss(o, t, w, bool(utw)) - float(g["loss_utw%d" % utw])) < 1e-6
    grad = LO.joints_mse_grad(o, t, w, bool(utw))
    assert np.abs(grad - g["grad_utw%d" % utw]).max() < 1e-9


def test_adam_three_steps(g):
    p, m, v = g["adam_p0"].copy(), np.zeros_like(g["adam_p0"]), np.zeros_like(g["adam_p0"])
    for i in range(3):
        p, m, v = LO.adam_step(p, g["adam_g%d" % i], m, v, i + 1)
        assert np.abs(p - g["adam_p%d" % (i + 1)]).max() < 2e-7
    assert np.abs(m - g["adam_m3"]).max() < 1e-7 and np.abs(v - g["adam_v3"]).max() < 1e-6


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_flip_back_live_reference(g):
    T = R._import()["transforms"]
    x = np.random.default_rng(5).standard_normal((2, 17, 6, 5)).astype(np.float32)
    pairs = g["flip_pairs"].tolist()
    assert np.array_equal(np.ascontiguousarray(T.flip_back(x.copy(), pairs)), LO.flip_back(x, pairs))


@pytest.fixture(scope="module")
def gt(golden_dir):
    return np.load(os.path.join(golden_dir, "train_ops.npz"))


@pytest.mark.parametrize("name", ["w48_384x288", "w32_256x192_s2", "odd_stride"])
def test_generate_target_vs_reference_golden(gt, name):
    """oracle restatement == the reference's JointsDataset.generate_target, bit for bit (clipped and outside stamps,
    invisible joints, per-joint weights)"""
    img_w, img_h, hm_w, hm_h, sigma = gt["tg_%s_meta" % name].tolist()
    jw = gt["tg_%s_jw" % name] if ("tg_%s_jw" % name) in gt.files else None
    j, v = gt["tg_%s_joints" % name], gt["tg_%s_vis" % name]
    for b in range(j.shape[0]):
        t, w = LO.generate_target(j[b], v[b], (img_w, img_h), (hm_w, hm_h), sigma, jw)
        assert np.array_equal(t, gt["tg_%s_target" % name][b]) and np.array_equal(w, gt["tg_%s_weight" % name][b])
    assert gt["tg_%s_weight" % name].min() == 0 and gt["tg_%s_target" % name].max() == 1.0


def test_accuracy_vs_reference_golden(gt):
    acc, avg, cnt, pred = LO.accuracy(gt["acc_output"], gt["acc_target"])
    assert np.array_equal(acc, gt["acc_acc"]) and avg == float(gt["acc_avg"]) and cnt == int(gt["acc_cnt"])
    assert np.array_equal(pred, gt["acc_pred"])
    acc0, avg0, cnt0, _ = LO.accuracy(gt["acc_output"][:2], np.zeros_like(gt["acc_target"][:2]))
    assert np.array_equal(acc0, gt["acc0_acc"]) and avg0 == float(gt["acc0_avg"]) and cnt0 == int(gt["acc0_cnt"]) == 0
