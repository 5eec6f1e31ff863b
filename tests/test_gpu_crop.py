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


def test_recrop_refinement_loop_matches_reference_composition():
    """Row f-2 + f-3 + a-R composed as the reference's dataset composes them at test time: the device-built model input
    of a refinement pass is bit-identical to cv2.warpAffine + ToTensor/Normalize + the rasterised, re-projected
    keypoints; the loop returns reference-format results whose next-pass boxes follow _load_coco_pose_results."""
    import json

    from buctd_b200 import function, pose_hrnet_coam, results
    from oracle import condition_oracle as CO
    from oracle import make_golden as G
    name = "coam_w32_128x96_stacked"
    cfg, _ = G.case_cfg(name)
    net = pose_hrnet_coam.get_pose_net(cfg, is_train=False)
    net.load_state_dict(G.case_state_dict(name), strict=True)
    net = net.cuda().eval()
    g = np.random.default_rng(11)
    imgs = {7: g.integers(0, 256, (200, 300, 3), dtype=np.uint8), 9: g.integers(0, 256, (240, 180, 3), dtype=np.uint8)}
    dev = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
    paths = {7: "data/coco/images/7.jpg", 9: "data/coco/images/9.jpg"}
    proposals = []
    for n, img_id in enumerate([7, 9, 7]):
        h, w = imgs[img_id].shape[:2]
        kp = np.zeros((17, 3))
        kp[:, 0], kp[:, 1], kp[:, 2] = g.uniform(0.2 * w, 0.8 * w, 17), g.uniform(0.2 * h, 0.8 * h, 17), g.uniform(0.3, 1, 17)
        proposals.append({"image_id": img_id, "keypoints": kp.reshape(-1).tolist(), "score": float(g.uniform(0.5, 1)), "annotation_id": 100 + n})
    shapes = {k: v.shape[:2] for k, v in imgs.items()}
    db = results.load_coco_pose_results(proposals, shapes, 17, 25, 96 * 1.0 / 128)
    ids = sorted(imgs)
    x, center, scale, trans = function.build_inputs_from_results(cfg, [dev[i] for i in ids], [ids.index(e["image_id"]) for e in db], db)
    got = x.cpu().numpy()
    for n, e in enumerate(db):
        M = WO.get_affine_transform(e["center"], e["scale"], 0, (96, 128), solver=__import__("cv2").getAffineTransform)
        assert np.array_equal(M, trans[n])
        rgb = WO.to_tensor_normalize(WO.warp_affine_u8(imgs[e["image_id"]], M, (96, 128)), MEAN, STD)
        assert np.array_equal(got[n, :3], rgb)
        kp = np.stack([M @ np.array([p[0], p[1], 1.0]) for p in e["cond_joints"]])
        assert np.abs(got[n, 3:] - CO.stacked_condition(kp, 128, 96).astype(np.float32)).max() <= 2e-4     # rasteriser tolerance
    hist = function.iterative_refine_recrop(cfg, net, dev, ids, paths, proposals, passes=2)
    assert len(hist) == 2 and all(len(h) == 3 for h in hist)
    json.dumps(hist)                                                          # plain Python numbers throughout
    p1, m1 = function.validate_batch(cfg, net, x, center, scale)
    first = {r["annotation_id"]: r for r in hist[0]}
    for n, r in enumerate(proposals):
        k = np.array(first[r["annotation_id"]]["keypoints"]).reshape(17, 3)
        # same input, same plan => the same bits (the forward is run-to-run deterministic, tests/test_gpu_determinism.py)
        assert np.array_equal(k[:, 2:].astype(np.float32), m1[n].cpu().numpy())
        assert np.array_equal(k[:, :2].astype(np.float32), p1[n].cpu().numpy())
    db2 = results.load_coco_pose_results(hist[0], shapes, 17, 25, 96 * 1.0 / 128)
    assert all(np.all(np.isfinite(e["center"])) and np.all(e["scale"] > 0) for e in db2)
