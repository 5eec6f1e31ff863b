"""The condition rasteriser restatement vs OpenCV itself (the arithmetic the reference calls:
cv2.GaussianBlur(ksize=(15,15), sigma=0), lib/dataset/JointsDataset.py:457-463)."""
import numpy as np
import pytest

from oracle import condition_oracle as C

cv2 = pytest.importorskip("cv2")


def test_blur_matches_cv2():
    g = np.random.default_rng(0)
    h, w = 96, 72
    for y, x in [(0, 0), (h - 1, w - 1), (3, w - 2), (h - 2, 5), (7, 7), (40, 30)]:
        m = np.zeros((h, w)); m[y, x] = 255
        assert np.abs(cv2.GaussianBlur(m, (15, 15), 0) - C.gaussian_blur_15(m)).max() < 1e-10
    m = g.uniform(size=(40, 30, 3))
    assert np.abs(cv2.GaussianBlur(m, (15, 15), 0) - C.gaussian_blur_15(m)).max() < 1e-12


def _ref_stacked(kpts, h, w):
    # literal restatement with cv2 of get_stacked_condition for cross-checking the numpy path
    out = []
    for k in np.array(kpts).astype(int):
        m = np.zeros((h, w))
        if 0 < k[0] < w and 0 < k[1] < h:
            m[k[1] - 1][k[0] - 1] = 255
        m = cv2.GaussianBlur(m, (15, 15), 0)
        if m.max() > 0:
            m /= m.max() / 255
        out.append(m)
    return np.array(out)


def test_stacked_condition_and_skips():
    kp = np.array([[10.7, 20.2], [0, 0], [0.9, 5], [71.9, 95.9], [72, 50], [30, 96], [-3, 4], [1, 1]])
    a = C.stacked_condition(kp, 96, 72)
    assert np.abs(a - _ref_stacked(kp, 96, 72)).max() < 1e-9
    assert a[1].max() == 0 and a[2].max() == 0 and a[4].max() == 0 and a[5].max() == 0 and a[6].max() == 0
    assert abs(a[0].max() - 255) < 1e-9 and a[0].argmax() == (20 - 1) * 72 + (10 - 1)
    assert a[7].argmax() == 0


def test_colored_and_plain():
    kp = np.array([[10, 20], [40, 60], [40, 60]])
    cols = C.default_colors(3)
    c = C.colored_condition(kp, 96, 72, cols)
    assert c.shape == (3, 96, 72) and abs(c.max() - 255) < 1e-9
    p = C.plain_condition(kp, 96, 72)
    assert p.dtype.kind == "i" and p.shape == (3, 96, 72) and (p[0] == p[2]).all()


@pytest.mark.parametrize("name", ["s128x96", "s384x288", "s40x30", "all_skipped", "single"])
def test_oracle_matches_reference_rasteriser_source(golden_dir, name):
    """tests/golden/condition.npz holds the outputs of the reference's OWN get_stacked_condition / get_condition_image /
    get_condition_image_colored / generate_heatmap source (lib/dataset/JointsDataset.py:457-543, compiled in isolation by
    oracle/make_golden_condition.py): the restatement must reproduce them, incl. borders, skips and duplicates."""
    import os
    g = np.load(os.path.join(golden_dir, "condition.npz"))
    kp = g[name + ":kpts"]
    h, w = [int(v) for v in g[name + ":size"]]
    cols = [list(g["colors"][i % len(g["colors"])]) for i in range(len(kp))]
    assert np.abs(C.stacked_condition(kp, h, w) - g[name + ":stacked"]).max() <= 3e-5        # the fixture is float32
    assert np.abs(C.colored_condition(kp, h, w, cols) - g[name + ":colored"]).max() <= 3e-5
    assert np.array_equal(C.plain_condition(kp, h, w), g[name + ":plain"])
    assert g["colors"].tolist() == C.COCO_KPT_COLORS


@pytest.mark.skipif(not __import__("oracle.refharness", fromlist=["x"]).available(), reason="reference tree not present (GPU box)")
def test_condition_golden_is_current(golden_dir):
    import os
    from oracle import make_golden_condition as M
    ref = M.reference_rasteriser()
    g = np.load(os.path.join(golden_dir, "condition.npz"))
    h, w, kp = M.cases()["s40x30"]
    st = np.moveaxis(np.asarray(ref.get_stacked_condition(kp.copy(), (h, w), None)), -1, 0)
    assert np.abs(st - g["s40x30:stacked"]).max() <= 3e-5
