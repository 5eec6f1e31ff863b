"""Device-side mirror of the hot section of the reference's validation loop (``validate``,
/root/reference/lib/core/function.py:200-270): forward, flip test (input flip -> forward -> flip_back ->
SHIFT_HEATMAP -> average) and ``get_final_preds``, with nothing leaving the GPU until the decoded keypoints
(204 B per crop) -- and the synthetic fixed-crop iterative refinement of BASELINE.json configs[3]
(SURVEY.md 8d "Config 4": pass n+1 is conditioned on the rasterised decode of pass n)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

import numpy as np

from . import condition, crop, inference, results, transforms

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]   # lib/dataset/coco.py:66-67


def fliplr_joints(joints, joints_vis, width, matched_parts):
    """``utils.transforms.fliplr_joints`` (lib/utils/transforms.py:60-74) for a batch [B, K, 3] (numpy, new arrays)"""
    j, v = np.array(joints, dtype=np.float64, copy=True), np.array(joints_vis, dtype=np.float64, copy=True)
    j[..., 0] = width - j[..., 0] - 1
    for p0, p1 in matched_parts:
        j[..., [p0, p1], :] = j[..., [p1, p0], :]
        v[..., [p0, p1], :] = v[..., [p1, p0], :]
    return j * v, v


@torch.no_grad()
def forward_flip_test(config, model, input: torch.Tensor, flip_pairs: Sequence[Sequence[int]] = COCO_FLIP_PAIRS,
                      cond_joints=None, cond_joints_vis=None, colors=None) -> torch.Tensor:
    """``output`` of validate() after the FLIP_TEST block (function.py:202-236); reads TEST.FLIP_TEST,
    TEST.SHIFT_HEATMAP, MODEL.CONDITIONAL_TOPDOWN.  The flipped condition follows ``flip_hm`` (transforms.py:33-57), which
    dispatches on the CHANNEL COUNT of the condition: > 3 channels = stacked (mirror + swap the planes of the flip
    pairs); exactly 3 = coloured / plain image, re-rasterised from the mirrored proposals ``meta['cond_joints']``
    (``cond_joints`` [B, K, 3], ``cond_joints_vis``) -- on the device here; fewer = mirrored only."""
    x = input
    if not config.MODEL.CONDITIONAL_TOPDOWN:
        x = x[:, :3]                                           # function.py:202-203
    x = x.cuda()
    output = model(x)
    if not config.TEST.FLIP_TEST:
        return output
    if config.MODEL.CONDITIONAL_TOPDOWN and x.shape[1] > 3:
        nc = x.shape[1] - 3
        if nc == 3:
            if cond_joints is None or cond_joints_vis is None:
                raise ValueError("a 3-channel condition is re-rasterised from the mirrored proposals: pass meta['cond_joints'] and "
                                 "meta['cond_joints_vis'] (lib/utils/transforms.py:38-47)")
            cj = cond_joints.cpu().numpy() if torch.is_tensor(cond_joints) else np.asarray(cond_joints)
            cv = cond_joints_vis.cpu().numpy() if torch.is_tensor(cond_joints_vis) else np.asarray(cond_joints_vis)
            W = int(config.MODEL.IMAGE_SIZE[0])
            fj, _ = fliplr_joints(cj, cv, W, flip_pairs)
            x_f = transforms.flip_channels_device(x, list(range(x.shape[1])))          # RGB planes mirrored (condition overwritten next)
            # the reference always calls get_condition_image_colored here, whatever DATASET.COLORED says
            condition.rasterize_condition(torch.from_numpy(np.trunc(fj).astype(np.float32)).to(x.device), x.shape[2], x.shape[3], "colored",
                                          colors=colors, out=x_f, c_off=3)
        else:
            x_f = transforms.flip_input_device(x, flip_pairs, stacked=nc > 3)
    else:
        x_f = transforms.flip_channels_device(x, list(range(x.shape[1])))
    output_flipped = model(x_f)
    return transforms.flip_merge_device(output, output_flipped, flip_pairs, bool(config.TEST.SHIFT_HEATMAP))


@torch.no_grad()
def validate_batch(config, model, input, center, scale, flip_pairs=COCO_FLIP_PAIRS):
    """forward (+ flip test) + get_final_preds on the device: returns (preds [B,K,2], maxvals [B,K,1]) CUDA tensors"""
    output = forward_flip_test(config, model, input, flip_pairs)
    return inference.decode_device(output, center, scale, post_process=bool(config.TEST.POST_PROCESS))


_COPY_STREAMS = {}          # device index -> the side stream of predict_stream


@torch.no_grad()
def predict_stream(config, model, batches, post=None):
    """The evaluation loop's hot section over a stream of HOST batches, the way validate() walks its DataLoader
    (lib/core/function.py:200-270) -- but what crosses PCIe per crop is the normalised RGB crop and 204 B of proposals,
    not 17 rasterised float32 planes: ``batches`` yields (rgb [B,3,H,W] float32/float16 host tensor -- pinned for an
    asynchronous copy --, kpts [B,K,>=2] float32 crop pixels, center [B,2], scale [B,2]); every batch goes through
    ``model.predict_from_proposals`` (forward + ``get_final_preds``, head fused into the decode) and comes back as one pinned host tensor
    [B, K, 3] = (x, y, maxval).  The host -> device copy of batch i+1 runs on a side stream under the forward of batch i
    (two device slots); results are yielded one batch late so that the device -> host read never stalls the queue.
    ``post`` (optional) maps the device result [B, K, 3] before it is read back -- the multi-GPU loop passes the ONE
    collective there (distributed.gather_keypoints).  TEST.FLIP_TEST is not applied here (use validate_batch)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    main = torch.cuda.current_stream(dev)
    # ONE copy stream per device for the life of the process: the two device slots are allocated under it, and PyTorch's
    # caching allocator only hands a freed block back to the stream it was allocated on -- with a fresh stream per call
    # every call paid two 85 MB cudaMallocs (~0.25 s: tools/e2e_probe.py, a whole 50-step pass read 5 ms per step slower)
    side = _COPY_STREAMS.get(dev.index)
    if side is None:
        side = _COPY_STREAMS[dev.index] = torch.cuda.Stream(dev)
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    outs = [None, None]
    pp = bool(config.TEST.POST_PROCESS)

    def stage(i, batch):
        b = i & 1
        rgb, kp, c, s = batch
        with torch.cuda.stream(side):
            if i >= 2:
                side.wait_event(freed[b])
            if slots[b] is None or slots[b][0].shape != rgb.shape or slots[b][0].dtype != rgb.dtype or slots[b][1].shape != kp.shape:
                slots[b] = (torch.empty(rgb.shape, dtype=rgb.dtype, device=dev), torch.empty(kp.shape, dtype=torch.float32, device=dev),
                            torch.empty(c.shape, dtype=torch.float32, device=dev), torch.empty(s.shape, dtype=torch.float32, device=dev))
            for d, h in zip(slots[b], (rgb, kp, c, s)):
                d.copy_(torch.as_tensor(h), non_blocking=True)
            ready[b].record(side)

    it = iter(batches)
    try:
        nxt = next(it)
    except StopIteration:
        return
    stage(0, nxt)
    i = 0
    pending = None
    while nxt is not None:
        b = i & 1
        main.wait_event(ready[b])
        try:
            nxt = next(it)
            stage(i + 1, nxt)
        except StopIteration:
            nxt = None
        rgb_d, kp_d, c_d, s_d = slots[b]
        preds, maxvals = model.predict_from_proposals(rgb_d, kp_d, c_d, s_d, post_process=pp)
        res = torch.cat([preds, maxvals], 2)
        if post is not None:
            res = post(res)
        freed[b].record(main)
        if outs[b] is None or outs[b].shape != res.shape:
            outs[b] = torch.empty(res.shape, dtype=torch.float32).pin_memory()
        outs[b].copy_(res, non_blocking=True)
        done[b].record(main)
        if pending is not None:
            done[pending].synchronize()
            yield outs[pending]
        pending = b
        i += 1
    if pending is not None:
        done[pending].synchronize()
        yield outs[pending]


@torch.no_grad()
def iterative_refine(config, model, input: torch.Tensor, passes: int = 3, center=None, scale=None,
                     cond_mode: str = "stacked", return_all: bool = False):
    """BASELINE.json configs[3]: ``passes`` forwards; pass n+1's condition = rasterised keypoints decoded from pass n
    (crop fixed: kpt_crop = 4 * (x_hm, y_hm) with the quarter-pixel offset; the reference re-crops from a results
    JSON between passes).  The condition planes of ``input`` are overwritten in place on a device copy.
    Returns the final (preds, maxvals) -- in image coordinates if center/scale are given -- or every pass's."""
    x = input.cuda().contiguous().float().clone()
    H, W = x.shape[2], x.shape[3]
    outs: List[tuple] = []
    for n in range(passes):
        hm = model(x)
        stride = float(W) / float(hm.shape[3])
        coords, maxvals = inference.decode_device(hm, None, None, post_process=bool(config.TEST.POST_PROCESS))
        if center is not None:
            outs.append(inference.decode_device(hm, center, scale, post_process=bool(config.TEST.POST_PROCESS)))
        else:
            outs.append((coords, maxvals))
        if n + 1 < passes:
            kp = (coords * stride).contiguous()
            condition.rasterize_condition(kp, H, W, cond_mode, out=x, c_off=3)
    return outs if return_all else outs[-1]


@torch.no_grad()
def iterative_refine_from_proposals(config, model, rgb: torch.Tensor, kpts: torch.Tensor, passes: int = 3, center=None, scale=None,
                                    return_all: bool = False):
    """``iterative_refine`` on the fast-path entry: pass 1 is conditioned on the bottom-up proposals ``kpts`` [B,K,>=2]
    (crop pixels), pass n+1 on the keypoints decoded from pass n (kpt_crop = stride * (x_hm, y_hm), crop fixed); the
    condition planes never exist at full resolution.  rgb [B,3,H,W] float32/float16, host or device."""
    rgb = rgb.cuda(non_blocking=True)
    kp = kpts.cuda(non_blocking=True).float()[:, :, :2].contiguous()
    W = rgb.shape[3]
    outs: List[tuple] = []
    pp = bool(config.TEST.POST_PROCESS)
    for n in range(passes):
        hm = model.forward_from_proposals(rgb, kp)
        stride = float(W) / float(hm.shape[3])
        last = n + 1 == passes
        if center is not None and (last or return_all):
            outs.append(inference.decode_device(hm, center, scale, post_process=pp))
        if not last or center is None:
            coords, maxvals = inference.decode_device(hm, None, None, post_process=pp)
            if center is None:
                outs.append((coords, maxvals))
            kp = (coords * stride).contiguous()
    return outs if return_all else outs[-1]


@torch.no_grad()
def build_inputs_from_results(config, images, image_index, db, cond_mode: str = "stacked"):
    """The model input of one refinement pass, built on the device from the previous pass's results the way the
    reference's dataset does (lib/dataset/JointsDataset.py:217-288, 314-326 at test time: no flip, rot 0):
    crop matrix from (center, scale) -> cv2.warpAffine + ToTensor + Normalize for the RGB planes, the previous
    keypoints mapped into the crop (affine_transform) and rasterised for the condition planes.
    images: list of CUDA uint8 [h, w, 3] RGB images; image_index[n] = which image person n lives in;
    db: ``results.load_coco_pose_results`` entries.  Returns (x [N, 3 + Cc, H, W], center [N, 2], scale [N, 2], trans)."""
    W, H = int(config.MODEL.IMAGE_SIZE[0]), int(config.MODEL.IMAGE_SIZE[1])
    N = len(db)
    K = db[0]["cond_joints"].shape[0] if N else int(config.MODEL.NUM_JOINTS)
    cc = K if cond_mode == "stacked" else 3
    dev = images[0].device
    trans = np.zeros((N, 2, 3))
    kp = np.zeros((N, K, 3), np.float32)
    for n, e in enumerate(db):
        trans[n] = crop.get_affine_transform(e["center"], e["scale"], 0, (W, H))
        for i in range(K):
            if e["cond_joints_vis"][i, 0] > 0.0:
                # the rasteriser truncates toward zero like the reference's int(); truncating the float64 value here
                # keeps that exact through the float32 hand-over
                kp[n, i, 0:2] = np.trunc(crop.affine_transform(e["cond_joints"][i, 0:2], trans[n]))
            else:
                kp[n, i, 0:2] = np.trunc(e["cond_joints"][i, 0:2])
    x = torch.zeros(N, 3 + cc, H, W, dtype=torch.float32, device=dev)
    crop.warp_crops([images[image_index[n]] for n in range(N)], trans, (W, H), out=x, c_off=0)
    condition.rasterize_condition(torch.from_numpy(kp).to(dev), H, W, cond_mode, out=x, c_off=3)
    center = torch.from_numpy(np.stack([e["center"] for e in db]).astype(np.float32)).to(dev) if N else torch.zeros(0, 2, device=dev)
    scale = torch.from_numpy(np.stack([e["scale"] for e in db]).astype(np.float32)).to(dev) if N else torch.zeros(0, 2, device=dev)
    return x, center, scale, trans


@torch.no_grad()
def iterative_refine_recrop(config, model, images, image_ids, image_paths, proposals, passes: int = 3, cond_mode: str = "stacked",
                            flip_pairs=COCO_FLIP_PAIRS, bu_bbox_margin=None, in_vis_thre=None):
    """BASELINE.json configs[3] with the REFERENCE's semantics (SURVEY.md 8d "Config 4"): between passes the results go
    through the COCO results list (rescored), the person box is re-derived from the new keypoints
    (``_load_coco_pose_results``: min / max of the non-zero keypoints +- BU_BBOX_MARGIN, clipped, padded to the aspect
    ratio, x SCALE_THRE) and the image is re-cropped -- the reference does this offline through a JSON file
    (scripts/test/test_BUCTD_COAM_gen_sample.sh:21); here the list stays in memory and the crops never leave the GPU.
    images: dict image_id -> CUDA uint8 [h, w, 3] RGB; proposals: a results list (``image_id``, ``keypoints`` [K*3],
    ``score``) from the bottom-up model.  Returns the results list of every pass."""
    K = int(config.MODEL.NUM_JOINTS)
    W, H = int(config.MODEL.IMAGE_SIZE[0]), int(config.MODEL.IMAGE_SIZE[1])
    margin = config.DATASET.BU_BBOX_MARGIN if bu_bbox_margin is None else bu_bbox_margin
    thre = (config.TEST.get("IN_VIS_THRE", 0.0) if hasattr(config.TEST, "get") else getattr(config.TEST, "IN_VIS_THRE", 0.0)) \
        if in_vis_thre is None else in_vis_thre
    ids = sorted(images)
    imgs = [images[i] for i in ids]
    shapes = {i: (int(images[i].shape[0]), int(images[i].shape[1])) for i in ids}
    scale_thre = getattr(config.TEST, "SCALE_THRE", 1.25)
    cur, history = proposals, []
    for _ in range(passes):
        db = results.load_coco_pose_results(cur, shapes, K, margin, W * 1.0 / H, 200, scale_thre)
        index = [ids.index(e["image_id"]) for e in db]
        x, center, scale, _ = build_inputs_from_results(config, imgs, index, db, cond_mode)
        preds, maxvals = validate_batch(config, model, x, center, scale, flip_pairs)
        all_preds = torch.cat([preds, maxvals], 2).cpu().numpy()
        c, s = center.cpu().numpy(), scale.cpu().numpy()
        boxes = np.zeros((len(db), 7))
        boxes[:, 0:2], boxes[:, 2:4] = c, s
        boxes[:, 4] = np.prod(s * 200, 1)
        boxes[:, 5] = [e["score"] for e in db]
        boxes[:, 6] = [r.get("annotation_id", -1) for r in cur]
        cur = results.coco_keypoint_results(all_preds, boxes, [e["image_id"] for e in db],
                                            [image_paths[e["image_id"]] for e in db], thre)
        history.append(cur)
    return history
