"""not gpu: oracle/train_oracle.py (training-mode forward + autograd, row a-T) against tests/golden/train_step_*.npz --
loss, heat-maps, every parameter's gradient norm and leading entries, and the updated BatchNorm running statistics of the
UNMODIFIED reference in train() mode (oracle/make_golden_trainstep.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_trainstep as G
from oracle import train_oracle as T


@pytest.mark.parametrize("name", list(G.CASES))
def test_train_oracle_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg, _ = G.case_cfg(name)
    sd = G.case_state_dict(name)
    x, tgt, tw = [torch.from_numpy(a) for a in G.case_data(name)]
    loss, hm, grads, running = T.train_forward_backward(sd, cfg, x, tgt, tw, drop_p=0.0)
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(hm.numpy() - g["heatmaps"]).max() < 1e-4
    names = [str(n) for n in g["grad_names"]]
    assert sorted(names) == sorted(grads)
    for i, k in enumerate(names):
        gn = float(grads[k].double().norm())
        assert abs(gn - g["grad_norms"][i]) <= 2e-3 * g["grad_norms"][i] + 1e-9, (k, gn, g["grad_norms"][i])
        head = grads[k].reshape(-1)[:12].numpy()
        assert np.abs(head - g["grad_heads"][i][:head.size]).max() <= 2e-3 * max(np.abs(g["grad_heads"][i]).max(), 1e-7) + 1e-8, k
    for key in g.files:
        if key.startswith("rm:"):
            assert np.abs(running[key[3:]][0].numpy() - g[key]).max() < 1e-5
        if key.startswith("rv:"):
            assert np.abs(running[key[3:]][1].numpy() - g[key]).max() < 1e-5


def test_dropout_mask_is_a_fair_hash():
    m = T.drop_mask(123, 0, 512, 512, 0.1)
    assert 0.89 < m.mean() < 0.91
    assert not np.array_equal(m, T.drop_mask(124, 0, 512, 512, 0.1))
    assert not np.array_equal(m, T.drop_mask(123, 1, 512, 512, 0.1))
    assert T.drop_mask(5, 0, 8, 8, 0.0).all()
