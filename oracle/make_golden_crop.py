"""TEST INFRASTRUCTURE -- writes tests/golden/crop_warp.npz: a small 8-bit image, crop matrices from the reference's
get_affine_transform (lib/utils/transforms.py:86-118), the crops cv2.warpAffine makes of it
(lib/dataset/JointsDataset.py:283-287) and torchvision's ToTensor + Normalize of their RGB view (tools/train.py:171-177).
    python -m oracle.make_golden_crop"""
import os
import sys

import cv2
import numpy as np
import torch
import torchvision.transforms as transforms

from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    if not R.available():
        sys.exit("reference tree not found at %s" % R.REF_ROOT)
    T = R._import()["transforms"]
    g = np.random.default_rng(67)
    img = cv2.GaussianBlur(g.integers(0, 256, (90, 120, 3), dtype=np.uint8), (5, 5), 0)
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    trans, u8, f32 = [], [], []
    for c, s, r in (((60, 45), 0.35, 0), ((10, 80), 0.5, 0), ((100, 20), 0.2, 25.0), ((60, 45), 0.9, -40.0)):
        M = T.get_affine_transform(np.array(c, np.float32), np.array([s * 0.75, s], np.float32), r, (48, 64))
        crop = cv2.warpAffine(img, M, (48, 64), flags=cv2.INTER_LINEAR)
        trans.append(M)
        u8.append(crop)
        f32.append(tf(cv2.cvtColor(crop, cv2.COLOR_BGR2RGB)).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "crop_warp.npz"), image=img, trans=np.stack(trans), crop_u8=np.stack(u8),
                        crop_f32=np.stack(f32))
    print("wrote crop_warp.npz")


if __name__ == "__main__":
    main()
