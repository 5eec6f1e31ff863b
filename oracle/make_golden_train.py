"""TEST INFRASTRUCTURE -- writes tests/golden/train_ops.npz from the UNMODIFIED reference:

* ``JointsDataset.generate_target`` (lib/dataset/JointsDataset.py:397-453).  The module cannot be imported here
  (matplotlib / pycocotools are absent), so the method's own source is compiled in isolation (ast) and called with a
  stand-in ``self`` that carries the five attributes it reads;
* ``core.evaluate.accuracy`` (lib/core/evaluate.py:40-70), imported as is.

Run in the build container:   python -m oracle.make_golden_train
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types

import numpy as np

from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_generate_target():
    path = os.path.join(R.REF_ROOT, "lib", "dataset", "JointsDataset.py")
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "generate_target":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns["generate_target"]
    raise RuntimeError("generate_target not found in %s" % path)


def fake_self(image_size, heatmap_size, sigma, K, joints_weight=None):
    return types.SimpleNamespace(num_joints=K, target_type="gaussian", heatmap_size=np.array(heatmap_size),
                                 image_size=np.array(image_size), sigma=sigma,
                                 use_different_joints_weight=joints_weight is not None, joints_weight=joints_weight)


def target_cases():
    """(name, image_size (w, h), heatmap_size (w, h), sigma, joints [B,K,3] f64, joints_vis [B,K,3] f64, joints_weight)"""
    g = np.random.default_rng(41)
    cases = []
    for name, img, hm, sigma, B, K, jw in (("w48_384x288", (288, 384), (72, 96), 3, 6, 17, False),
                                           ("w32_256x192_s2", (192, 256), (48, 64), 2, 4, 17, True),
                                           ("odd_stride", (100, 90), (30, 20), 3, 3, 5, False)):
        j = np.zeros((B, K, 3))
        j[:, :, 0] = g.uniform(-40, img[0] + 40, (B, K))          # some stamps clipped, some completely outside
        j[:, :, 1] = g.uniform(-40, img[1] + 40, (B, K))
        j[0, 0, :2] = (0.0, 0.0)
        j[0, 1, :2] = (img[0] - 1e-9, img[1] - 1e-9)
        j[0, 2, :2] = (-13.0 * img[0] / hm[0], 5.0)               # exactly at the "outside" boundary
        j[0, 3, :2] = (img[0] + 9.49 * img[0] / hm[0], 7.0)
        v = np.zeros((B, K, 3))
        v[:, :, 0] = (g.uniform(size=(B, K)) < 0.8).astype(np.float64)
        v[:, :, 1] = v[:, :, 0]
        w = g.uniform(1.0, 1.5, (K, 1)).astype(np.float32) if jw else None
        cases.append((name, img, hm, sigma, j, v, w))
    return cases


def main():
    if not R.available():
        sys.exit("reference tree not found at %s" % R.REF_ROOT)
    gen = reference_generate_target()
    out = {}
    for name, img, hm, sigma, j, v, w in target_cases():
        self = fake_self(img, hm, sigma, j.shape[1], w)
        tg, tw = zip(*[gen(self, j[b], v[b]) for b in range(j.shape[0])])
        out["tg_%s_joints" % name], out["tg_%s_vis" % name] = j, v
        out["tg_%s_target" % name], out["tg_%s_weight" % name] = np.stack(tg), np.stack(tw)
        out["tg_%s_meta" % name] = np.array([img[0], img[1], hm[0], hm[1], sigma], np.int32)
        if w is not None:
            out["tg_%s_jw" % name] = w
    # accuracy(): outputs near / far from the targets, unlabeled joints (target arg-max at the origin)
    R._import()
    evaluate = importlib.import_module("core.evaluate")
    g = np.random.default_rng(43)
    name, img, hm, sigma, j, v, w = target_cases()[0]
    target = out["tg_%s_target" % name][:3]
    output = target + g.standard_normal(target.shape).astype(np.float32) * 0.35
    output[1] = np.roll(target[1], 4, axis=-1)                   # a crop whose predictions are all 4 px off
    acc, avg, cnt, pred = evaluate.accuracy(output, target)
    out["acc_output"], out["acc_target"] = output, target
    out["acc_acc"], out["acc_avg"], out["acc_cnt"], out["acc_pred"] = acc, np.float64(avg), np.int32(cnt), pred
    empty = np.zeros_like(target[:2])
    acc0, avg0, cnt0, _ = evaluate.accuracy(output[:2], empty)
    out["acc0_acc"], out["acc0_avg"], out["acc0_cnt"] = acc0, np.float64(avg0), np.int32(cnt0)
    np.savez_compressed(os.path.join(GOLDEN, "train_ops.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "train_ops.npz"), "acc", acc[:4], "avg", avg, "cnt", cnt)


if __name__ == "__main__":
    main()
