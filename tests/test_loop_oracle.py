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
