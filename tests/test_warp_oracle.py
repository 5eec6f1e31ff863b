"""not gpu: pins oracle/warp_oracle.py (crop pipeline) against OpenCV itself -- cv2.warpAffine and
cv2.getAffineTransform are the third-party arithmetic the reference calls (lib/dataset/JointsDataset.py:283-287,
lib/utils/transforms.py:113-116) -- and, where /root/reference is mounted, against the reference's
get_affine_transform."""
import numpy as np
import pytest

from oracle import refharness as R
from oracle import warp_oracle as WO

cv2 = pytest.importorskip("cv2")


def _cases(n, g):
    for t in range(n):
        f = g.uniform(0.3, 2.5)
        tx, ty = g.uniform(-200, 200, 2)
        th = 0.0 if t % 2 == 0 else g.uniform(-0.7, 0.7)
        yield np.array([[f * np.cos(th), -f * np.sin(th), tx], [f * np.sin(th), f * np.cos(th), ty]])


def test_warp_affine_u8_is_cv2_bit_for_bit():
    g = np.random.default_rng(0)
    img = g.integers(0, 256, (211, 317, 3), dtype=np.uint8)
    for M in _cases(10, g):
        assert np.array_equal(cv2.warpAffine(img, M, (96, 128), flags=cv2.INTER_LINEAR), WO.warp_affine_u8(img, M, (96, 128)))


def test_crop_matrices_and_crops_like_the_reference_pipeline():
    g = np.random.default_rng(1)
    img = g.integers(0, 256, (427, 640, 3), dtype=np.uint8)
    for _ in range(6):
        c = np.array([g.uniform(100, 540), g.uniform(80, 340)], np.float32)
        s = np.array([g.uniform(0.6, 2.4)] * 2, np.float32) * np.array([0.75, 1.0], np.float32)
        r = g.choice([0.0, 0.0, 17.5, -33.0])
        M_cv = WO.get_affine_transform(c, s, r, (288, 384), solver=cv2.getAffineTransform)
        M_np = WO.get_affine_transform(c, s, r, (288, 384))
        assert np.abs(M_cv - M_np).max() < 1e-9
        assert np.array_equal(cv2.warpAffine(img, M_cv, (288, 384), flags=cv2.INTER_LINEAR), WO.warp_affine_u8(img, M_cv, (288, 384)))


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_get_affine_transform_live_reference():
    T = R._import()["transforms"]
    g = np.random.default_rng(2)
    for inv in (0, 1):
        c = np.array([g.uniform(100, 540), g.uniform(80, 340)], np.float32)
        s = np.array([1.3, 1.7], np.float32)
        ref = T.get_affine_transform(c, s, 30.0, (288, 384), inv=inv)
        assert np.array_equal(ref, WO.get_affine_transform(c, s, 30.0, (288, 384), inv=inv, solver=cv2.getAffineTransform))


def test_golden_crop(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "crop_warp.npz"))
    for i in range(g["trans"].shape[0]):
        u8 = WO.warp_affine_u8(g["image"], g["trans"][i], (48, 64))
        assert np.array_equal(u8, g["crop_u8"][i])
        assert np.array_equal(WO.to_tensor_normalize(u8[:, :, ::-1], (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)), g["crop_f32"][i])
