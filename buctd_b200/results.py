"""Host-side mirror of the results round trip that closes BUCTD's refinement loop:

* ``rescore``                    the rescoring loop of ``evaluate``      /root/reference/lib/dataset/dataloader.py:596-612
* ``coco_keypoint_results``      ``_coco_keypoint_results_one_category_kernel`` + grouping  :564-590, :682-717
* ``write_coco_keypoint_results`` ``_write_coco_keypoint_results`` (json.dump sort_keys, indent 4)  :652-680
* ``xywh2cs``                    ``_xywh2cs``                            :304-321
* ``load_coco_pose_results``     ``_load_coco_pose_results`` (results JSON -> crops + conditions for the next pass) :454-508

This is file-format / bookkeeping code on a few hundred bytes per person: plain Python + NumPy on the host, as in the
reference.  Number types follow the reference's environment (numpy 1.22: ``0 + np.float32`` promotes to float64, so the
keypoint score is accumulated in float64)."""
from __future__ import annotations

import json
import os
from collections import OrderedDict

import numpy as np


def _valid_counts(all_preds, in_vis_thre):
    return (np.asarray(all_preds)[:, :, 2] > in_vis_thre).sum(1)


def rescore(all_preds, box_scores, in_vis_thre):
    """all_preds [N,K,3] (x, y, maxval), box_scores [N] -> (score, keypoint_score) float64 [N]:
    keypoint_score = mean of the maxvals above ``in_vis_thre`` (0 when none), score = keypoint_score * box_score."""
    v = np.asarray(all_preds)[:, :, 2]
    N, K = v.shape
    kpt = np.zeros(N, np.float64)
    cnt = np.zeros(N, np.int64)
    for j in range(K):                       # joint by joint: the reference's summation order
        t = v[:, j].astype(np.float64)
        m = v[:, j] > in_vis_thre
        kpt = np.where(m, kpt + t, kpt)
        cnt += m
    kpt = np.where(cnt != 0, kpt / np.maximum(cnt, 1), kpt)
    return kpt * np.asarray(box_scores, np.float64), kpt


def coco_keypoint_results(all_preds, all_boxes, image_ids, image_paths, in_vis_thre, cat_id=1):
    """The list ``evaluate`` writes (no OKS-NMS: with a bottom-up / json box source the reference discards it,
    lib/dataset/dataloader.py:627-630).  all_boxes [N,7] = (cx, cy, sx, sy, area, score, annotation_id) as filled by
    validate() (lib/core/function.py:266-270).  Persons are grouped per image in order of first appearance."""
    all_preds = np.asarray(all_preds)
    all_boxes = np.asarray(all_boxes, np.float64)
    score, kpt_score = rescore(all_preds, all_boxes[:, 5], in_vis_thre)
    valid = _valid_counts(all_preds, in_vis_thre)
    groups = OrderedDict()
    for idx in range(all_preds.shape[0]):
        groups.setdefault(image_ids[idx], []).append(idx)
    results = []
    for img, idxs in groups.items():
        for idx in idxs:
            kp = np.zeros(all_preds.shape[1] * 3, dtype=np.float64)
            kp[0::3], kp[1::3], kp[2::3] = all_preds[idx, :, 0], all_preds[idx, :, 1], all_preds[idx, :, 2]
            results.append({
                "image_id": img,
                "image_path": os.path.join(*image_paths[idx].split("/")[-3:]),
                "category_id": cat_id,
                "keypoints": [float(x) for x in kp],
                "score": float(score[idx]),
                "center": [float(x) for x in all_boxes[idx, 0:2]],
                "scale": [float(x) for x in all_boxes[idx, 2:4]],
                "annotation_id": int(all_boxes[idx, 6]),
                "box_score": float(all_boxes[idx, 5]),
                # the reference's accumulator starts as the int 0 and stays one when no joint passes the threshold
                "keypoint_score": float(kpt_score[idx]) if valid[idx] != 0 else 0,
            })
    return results


def write_coco_keypoint_results(results, res_file):
    with open(res_file, "w") as f:
        json.dump(results, f, sort_keys=True, indent=4)


def xywh2cs(x, y, w, h, aspect_ratio, pixel_std=200, scale_thre=1.25):
    """box -> (center f32 [2], scale f32 [2]): pad to the crop's aspect ratio, /pixel_std, * SCALE_THRE."""
    center = np.zeros((2), dtype=np.float32)
    center[0] = x + w * 0.5
    center[1] = y + h * 0.5
    if w > aspect_ratio * h:
        h = w * 1.0 / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    scale = np.array([w * 1.0 / pixel_std, h * 1.0 / pixel_std], dtype=np.float32)
    if center[0] != -1:
        scale = scale * scale_thre
    return center, scale


def load_coco_pose_results(results, image_shapes, num_joints, bu_bbox_margin, aspect_ratio, pixel_std=200, scale_thre=1.25):
    """``results``: the list a previous pass wrote (or its file name); ``image_shapes``: image_id -> (height, width)
    (the reference reads each image to get them).  Returns the reference's kpt_db entries without the image itself:
    bbox from the non-zero predicted keypoints +- margin clipped to the image, (center, scale), score, cond_joints."""
    if isinstance(results, str):
        with open(results) as f:
            results = json.load(f)
    db = []
    for r in results:
        cond = np.array(r["keypoints"]).reshape((num_joints, 3))
        h_img, w_img = image_shapes[r["image_id"]]
        xs, ys = cond[:, 0][np.nonzero(cond[:, 0])], cond[:, 1][np.nonzero(cond[:, 1])]
        xmin = np.clip(np.min(xs) - bu_bbox_margin, 0, w_img)
        ymin = np.clip(np.min(ys) - bu_bbox_margin, 0, h_img)
        xmax = np.clip(np.max(xs) + bu_bbox_margin, 0, w_img)
        ymax = np.clip(np.max(ys) + bu_bbox_margin, 0, h_img)
        c, s = xywh2cs(xmin, ymin, xmax - xmin, ymax - ymin, aspect_ratio, pixel_std, scale_thre)
        db.append({"image_id": r["image_id"], "center": c, "scale": s, "score": r["score"],
                   "joints_3d": np.zeros((num_joints, 3)), "joints_3d_vis": np.ones((num_joints, 3)),
                   "cond_joints": cond, "cond_joints_vis": np.ones((num_joints, 3)),
                   "bbox": [xmin, ymin, xmax - xmin, ymax - ymin], "cond_max_iou": 1})
    return db
