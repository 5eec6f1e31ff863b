"""Pins oracle/ (the CPU restatement) against outputs of the reference's own code.

tests/golden/*.npz were written by oracle/make_golden.py from the UNMODIFIED reference modules
(/root/reference, fp32, CPU).  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so these fixtures are the pin.
"""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle as D
from oracle import hrnet_oracle as O
from oracle import make_golden as G
from oracle import refharness as R


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


# fp32 CPU convolutions are not bit-reproducible across oneDNN code paths / thread counts;
# 1e-4 is 10x below the tightest product tolerance (1e-3) and ~100x above observed differences.
ORACLE_TOL = 1e-4


@pytest.mark.parametrize("name", ["plain_w32_256x192", "coam_w32_128x96_stacked", "coam_w32_128x96_plaincond",
                                  "coam_w32_128x96_att1000", "coam_w32_128x96_att0011", "coam_w32_128x96_att0001"])
def test_forward_matches_reference_small(golden_dir, name):
    cfg, _ = G.case_cfg(name)
    g = _load(golden_dir, name)
    out = O.forward(G.case_state_dict(name), cfg, torch.from_numpy(G.case_inputs(name))).numpy()
    assert out.shape == g["heatmaps"].shape
    assert np.abs(out - g["heatmaps"]).max() < ORACLE_TOL


@pytest.mark.slow
@pytest.mark.parametrize("name", ["coam_w48_384x288_stacked", "coam_w48_384x288_colored", "prenet_w48_384x288",
                                  "coam_w48_384x288_att0011"])
def test_forward_matches_reference_full(golden_dir, name):
    cfg, _ = G.case_cfg(name)
    g = _load(golden_dir, name)
    out = O.forward(G.case_state_dict(name), cfg, torch.from_numpy(G.case_inputs(name))).numpy()
    assert np.abs(g["heatmaps"]).max() > 0.5          # fixture is not vacuous
    assert np.abs(out - g["heatmaps"]).max() < ORACLE_TOL


def test_condition_empty_raises():
    cfg, _ = G.case_cfg("coam_w32_128x96_stacked")
    x = torch.zeros(1, 3, 128, 96)
    with pytest.raises(Exception, match="condition is empty"):
        O.forward(G.case_state_dict("coam_w32_128x96_stacked"), cfg, x)


@pytest.mark.parametrize("name", list(G.CASES))
def test_decode_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    coords, maxvals = D.get_max_preds(g["heatmaps"])
    assert np.array_equal(coords, g["coords"])                     # integer indices: bit exact
    assert np.array_equal(maxvals, g["maxvals"])
    preds, mv = D.get_final_preds(True, g["heatmaps"], g["center"], g["scale"])
    assert np.array_equal(mv, g["maxvals"])
    assert np.abs(preds - g["preds"]).max() <= 1e-4                # f64 affine, cast to f32


def test_decode_edge_cases(golden_dir):
    g = _load(golden_dir, "decode_edge_cases")
    coords, maxvals = D.get_max_preds(g["heatmaps"])
    assert np.array_equal(coords, g["coords"])
    assert np.array_equal(maxvals, g["maxvals"])
    for pp in (0, 1):
        preds, _ = D.get_final_preds(bool(pp), g["heatmaps"], g["center"], g["scale"])
        assert np.abs(preds - g["preds_pp%d" % pp]).max() <= 1e-4
    # spot checks of the semantics the fixture encodes
    assert tuple(coords[0, 2]) == (7.0, 5.0)        # tie -> first flat index
    assert tuple(coords[0, 9]) == (0.0, 0.0) and maxvals[0, 9, 0] == np.float32(-0.1)


def test_joints_mse(golden_dir):
    g = _load(golden_dir, "joints_mse")
    for utw in (0, 1):
        v = D.joints_mse_loss(g["output"], g["target"], g["target_weight"], bool(utw))
        assert abs(v - float(g["loss_utw%d" % utw])) < 1e-6


@pytest.mark.skipif(not R.available(), reason="reference tree not present (GPU box)")
def test_live_reference_matches_golden(golden_dir):
    """Where the reference tree exists, re-run it and make sure the committed fixture is current."""
    name = "coam_w32_128x96_stacked"
    cfg, _ = G.case_cfg(name)
    net = R.build_model(cfg, G.case_state_dict(name))
    hm = R.forward(net, torch.from_numpy(G.case_inputs(name))).numpy()
    assert np.abs(hm - _load(golden_dir, name)["heatmaps"]).max() < ORACLE_TOL
