"""TEST INFRASTRUCTURE -- writes tests/golden/results_roundtrip.json from the UNMODIFIED reference methods
``evaluate`` (rescoring + grouping), ``_write_coco_keypoint_results``, ``_coco_keypoint_results_one_category_kernel``,
``_xywh2cs`` and ``_load_coco_pose_results`` of lib/dataset/dataloader.py (:304-321, :454-508, :538-717).

The module itself cannot be imported here (pycocotools, json_tricks, the compiled nms extension are absent), so the
five methods' own source is compiled in isolation (ast) and run against a stand-in ``self``: a two-image fake COCO
index, ``cv2.imread`` replaced by a function returning a blank image of the recorded size, the standard ``json``
module in place of json_tricks, ``np.float`` = float (numpy >= 1.24 dropped the alias) and the predictions passed
as float64 (= what numpy 1.22's ``0 + np.float32`` promotion makes of them in the reference's environment).
``is_train=True`` makes ``evaluate`` return right after writing the file (no pycocotools AP).

    python -m oracle.make_golden_results
"""
from __future__ import annotations

import ast
import json
import logging
import os
import sys
import tempfile
import types
from collections import OrderedDict, defaultdict

import numpy as np

from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WANTED = ("evaluate", "_write_coco_keypoint_results", "_coco_keypoint_results_one_category_kernel", "_xywh2cs",
          "_load_coco_pose_results")


def reference_methods(image_shapes):
    path = os.path.join(R.REF_ROOT, "lib", "dataset", "dataloader.py")
    tree = ast.parse(open(path).read())
    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    np_shim.float = float
    by_name = {}

    def imread(name, flags=None):
        h, w = image_shapes[os.path.basename(name)]
        return np.zeros((h, w, 3), np.uint8)

    cv2_shim = types.SimpleNamespace(imread=imread, IMREAD_COLOR=1, IMREAD_IGNORE_ORIENTATION=128)
    ns = {"np": np_shim, "os": os, "json": json, "defaultdict": defaultdict, "OrderedDict": OrderedDict, "cv2": cv2_shim,
          "logger": logging.getLogger("ref"), "oks_nms": lambda *a, **k: [], "soft_oks_nms": lambda *a, **k: []}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in WANTED and node.name not in by_name:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            by_name[node.name] = ns[node.name]
    assert set(by_name) == set(WANTED), set(WANTED) - set(by_name)
    return by_name


def inputs():
    g = np.random.default_rng(59)
    N, K = 7, 17
    img_of = [0, 1, 0, 2, 1, 0, 2]
    ids = [139, 785, 139, 872, 785, 139, 872]
    names = {139: "000000000139.jpg", 785: "000000000785.jpg", 872: "000000000872.jpg"}
    shapes = {"000000000139.jpg": (426, 640), "000000000785.jpg": (425, 640), "000000000872.jpg": (640, 621)}
    preds = np.zeros((N, K, 3), np.float32)
    preds[:, :, 0] = g.uniform(5, 600, (N, K))
    preds[:, :, 1] = g.uniform(5, 420, (N, K))
    preds[:, :, 2] = g.uniform(0.0, 1.0, (N, K))
    preds[3, :, 2] = 0.05                         # nothing above IN_VIS_THRE: keypoint_score stays 0
    preds[5, 4, 0] = 0.0                          # a zero coordinate is skipped when the next pass derives the box
    boxes = np.zeros((N, 7))
    boxes[:, 0:2] = g.uniform(100, 400, (N, 2))
    boxes[:, 2:4] = g.uniform(0.5, 2.5, (N, 2))
    boxes[:, 4] = np.prod(boxes[:, 2:4] * 200, 1)
    boxes[:, 5] = g.uniform(0.3, 1.0, N)
    boxes[:, 6] = np.arange(1000, 1000 + N)
    paths = ["data/coco/images/val2017/" + names[i] for i in ids]
    return preds, boxes, ids, paths, names, shapes


def main():
    if not R.available():
        sys.exit("reference tree not found at %s" % R.REF_ROOT)
    preds, boxes, ids, paths, names, shapes = inputs()
    m = reference_methods(shapes)
    tmp = tempfile.mkdtemp()
    gt_file = os.path.join(tmp, "gt.json")
    with open(gt_file, "w") as f:
        json.dump({"images": [{"id": i, "file_name": n} for i, n in names.items()]}, f)
    coco = types.SimpleNamespace(loadImgs=lambda index: [{"file_name": names[index], "id": index}],
                                 anns={1000 + i: {"id": 1000 + i, "area": 1.0} for i in range(len(ids))})
    self = types.SimpleNamespace(mode="val", image_set_index=list(names), coco=coco, image_dir="data/coco/images/val2017",
                                 is_train=True, use_gt_bbox=False, use_bu_bbox_test=True, use_bu_bbox_train=False,
                                 num_joints=17, in_vis_thre=0.2, oks_thre=0.9, soft_nms=False, joints_weight=np.ones((17, 1)),
                                 classes=["__background__", "person"], _class_to_coco_ind={"person": 1},
                                 aspect_ratio=288 * 1.0 / 384, pixel_std=200, scale_thre=1.25, bu_bbox_margin=25,
                                 img_dir="data/coco/images/val2017", test_gt_file=gt_file)
    for name, fn in m.items():
        setattr(self, name, types.MethodType(fn, self))
    cfg = types.SimpleNamespace(OUTPUT_JSON=os.path.join(tmp, "res.json"), TEST=types.SimpleNamespace(COCO_BBOX_FILE="x.json"))
    # areas[...] is only read on the is_train path: annotation ids index coco.anns
    self.evaluate(cfg, preds.astype(np.float64), tmp, boxes, paths)
    text = open(cfg.OUTPUT_JSON).read()
    self.bbox_file = cfg.OUTPUT_JSON
    db = self._load_coco_pose_results()
    out = {"results_json_text": text,
           "inputs": {"preds": preds.tolist(), "boxes": boxes.tolist(), "ids": ids, "paths": paths,
                      "shapes": {str(i): list(shapes[n]) for i, n in names.items()}},
           "next_pass": [{"center": [float(x) for x in e["center"]], "scale": [float(x) for x in e["scale"]], "score": e["score"],
                          "bbox": [float(x) for x in e["bbox"]], "cond_joints": np.asarray(e["cond_joints"]).tolist()} for e in db]}
    with open(os.path.join(GOLDEN, "results_roundtrip.json"), "w") as f:
        json.dump(out, f)
    print("wrote results_roundtrip.json:", len(db), "persons;", len(text), "bytes of results JSON")


if __name__ == "__main__":
    main()
