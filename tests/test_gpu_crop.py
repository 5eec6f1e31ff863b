"""-m gpu: crop warp + normalise on the device through the C ABI, bit for bit against the golden
crops cv2.warpAffine + torchvision made (tests/golden/crop_warp.npz) and against the numpy oracle at full size."""
import os

import numpy as np
import pytest
import torch

from buctd_b200 import crop
from oracle import warp_oracle as WO

pytestmark = pytest.mark.gpu
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def test_golden_crops_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "crop_warp.npz"))
    img = torch.from_numpy(g["image"]).cuda()
    B = g["trans"].shape[0]
    out, u8 = crop.warp_crops([img] * B, g["trans"], (48, 64), bgr_to_rgb=True, want_u8=True)
    assert np.array_equal(u8.cpu().numpy()[..., ::-1], g["crop_u8"])          # out_u8 is written in output (RGB) order
    assert np.array_equal(out.cpu().numpy(), g["crop_f32"])


def test_full_size_crops_into_model_input_vs_oracle():
    """384x288 crops of two differently sized images written into planes 0..2 of a 20-channel model input"""
    g = np.random.default_rng(5)
    imgs = [g.integers(0, 256, (427, 640, 3), dtype=np.uint8), g.integers(0, 256, (333, 500, 3), dtype=np.uint8)]
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    which = [0, 1, 1, 0, 0]
    trans = []
    for b in which:
        h, w = imgs[b].shape[:2]
        c = np.array([g.uniform(0.2 * w, 0.8 * w), g.uniform(0.2 * h, 0.8 * h)], np.float32)
        s = np.array([0.75, 1.0], np.float32) * np.float32(g.uniform(0.6, 2.6))
        trans.append(crop.get_affine_transform(c, s, float(g.choice([0.0, 0.0, 30.0])), (288, 384)))
    x = torch.full((len(which), 20, 384, 288), 7.0, device="cuda")
    crop.warp_crops([dev[b] for b in which], np.stack(trans), (288, 384), out=x, c_off=0)
    got = x.cpu().numpy()
    assert np.all(got[:, 3:] == 7.0)                                         # the condition planes are not touched
    for i, b in enumerate(which):
        ref = WO.to_tensor_normalize(WO.warp_affine_u8(imgs[b], trans[i], (288, 384)), MEAN, STD)
        assert np.array_equal(got[i, :3], ref)


def test_empty_batch_and_cpu_input():
    out = crop.warp_crops([], np.zeros((0, 2, 3)), (288, 384))
    assert tuple(out.shape) == (0, 3, 384, 288)
    with pytest.raises(Exception, match="no CPU path"):
        crop.warp_crops([torch.zeros(4, 4, 3, dtype=torch.uint8)], np.zeros((1, 2, 3)), (8, 8))
