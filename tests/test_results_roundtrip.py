"""not gpu: buctd_b200.results (rescoring, results JSON, next-pass loader: SURVEY.md 8 f-3) against
tests/golden/results_roundtrip.json, written by oracle/make_golden_results.py from the reference's own evaluate /
_write_coco_keypoint_results / _load_coco_pose_results source."""
import json
import os

import numpy as np
import pytest

from buctd_b200 import results as RS


@pytest.fixture(scope="module")
def g(golden_dir):
    with open(os.path.join(golden_dir, "results_roundtrip.json")) as f:
        return json.load(f)


def _inputs(g):
    i = g["inputs"]
    return np.array(i["preds"], np.float32), np.array(i["boxes"]), i["ids"], i["paths"]


def test_results_json_is_byte_identical(g, tmp_path):
    preds, boxes, ids, paths = _inputs(g)
    res = RS.coco_keypoint_results(preds, boxes, ids, paths, in_vis_thre=0.2)
    out = tmp_path / "res.json"
    RS.write_coco_keypoint_results(res, str(out))
    assert out.read_text() == g["results_json_text"]


def test_rescoring(g):
    preds, boxes, ids, paths = _inputs(g)
    score, kpt = RS.rescore(preds, boxes[:, 5], 0.2)
    assert kpt[3] == 0 and score[3] == 0                     # nothing above IN_VIS_THRE
    ref = {(r["annotation_id"]): r for r in json.loads(g["results_json_text"])}
    for n in range(len(ids)):
        r = ref[int(boxes[n, 6])]
        assert r["score"] == score[n] and r["keypoint_score"] == kpt[n] and r["box_score"] == boxes[n, 5]


def test_next_pass_crops_from_results(g):
    shapes = {int(k): tuple(v) for k, v in g["inputs"]["shapes"].items()}
    db = RS.load_coco_pose_results(json.loads(g["results_json_text"]), shapes, 17, 25, 288 * 1.0 / 384)
    assert len(db) == len(g["next_pass"])
    for e, r in zip(db, g["next_pass"]):
        assert [float(x) for x in e["center"]] == r["center"] and [float(x) for x in e["scale"]] == r["scale"]
        assert [float(x) for x in e["bbox"]] == r["bbox"] and e["score"] == r["score"]
        assert np.array_equal(e["cond_joints"], np.array(r["cond_joints"]))


def test_xywh2cs_aspect_padding():
    c, s = RS.xywh2cs(10, 20, 100, 50, 0.75)
    assert c.dtype == np.float32 and np.allclose(c, [60, 45])
    assert np.allclose(s, np.array([100 / 200, 100 / 0.75 / 200], np.float32) * 1.25)
