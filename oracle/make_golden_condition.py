"""TEST INFRASTRUCTURE -- writes tests/golden/condition.npz from the UNMODIFIED reference rasteriser (row a-R):

``JointsDataset.generate_heatmap``, ``get_stacked_condition``, ``get_condition_image``, ``get_condition_image_colored``
(lib/dataset/JointsDataset.py:457-543).  The module cannot be imported here (matplotlib / pycocotools are absent), so the
four methods' own source is compiled in isolation (ast) against ``np`` and ``cv2`` -- the only names they use -- and called
with a stand-in ``self`` whose ``generate_heatmap`` is the reference's.  The colour table is read out of
``lib/dataset/coco.py`` (the ``self.kpt_colors`` assignment, :54) the same way.

Run in the build container:   python -m oracle.make_golden_condition
"""
from __future__ import annotations

import ast
import os
import types

import numpy as np

from oracle import refharness as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
METHODS = ("generate_heatmap", "get_stacked_condition", "get_condition_image", "get_condition_image_colored")


def reference_rasteriser():
    """A stand-in ``self`` carrying the reference's four methods (bound)."""
    import cv2
    path = os.path.join(R.REF_ROOT, "lib", "dataset", "JointsDataset.py")
    tree = ast.parse(open(path).read())
    ns = {"np": np, "cv2": cv2, "os": os}
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in METHODS:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            found.append(node.name)
    assert sorted(found) == sorted(METHODS), found
    obj = types.SimpleNamespace()
    for m in METHODS:
        setattr(obj, m, types.MethodType(ns[m], obj))
    return obj


def reference_colors():
    """``self.kpt_colors`` of the COCO dataset (lib/dataset/coco.py), evaluated from its own source."""
    path = os.path.join(R.REF_ROOT, "lib", "dataset", "coco.py")
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Attribute) and t.attr == "kpt_colors" for t in node.targets):
            try:
                return [list(c) for c in ast.literal_eval(node.value)]
            except ValueError:
                continue
    raise RuntimeError("self.kpt_colors not found in %s" % path)


def cases():
    """name -> (h, w, kpts [K, 3]): interior, borders (reflect-101 folding), the strict 0 < x < W skips, (1, 1) -> pixel
    (0, 0), duplicates / overlapping blobs, fractional and negative coordinates (truncation toward zero)."""
    g = np.random.default_rng(5)
    out = {}
    for name, h, w in (("s128x96", 128, 96), ("s384x288", 384, 288), ("s40x30", 40, 30)):
        K = 17
        kp = np.zeros((K, 3))
        kp[:, 0] = g.uniform(1, w - 1, K)
        kp[:, 1] = g.uniform(1, h - 1, K)
        kp[0, :2] = (1, 1)                     # -> pixel (0, 0)
        kp[1, :2] = (w - 1, h - 1)             # last drawable pixel: (h-2, w-2)
        kp[2, :2] = (w, 10)                    # x == W: skipped
        kp[3, :2] = (10, h)                    # y == H: skipped
        kp[4, :2] = (0, 0)                     # "missing": skipped
        kp[5, :2] = (0.99, 7.5)                # truncates to x = 0: skipped
        kp[6, :2] = (-3.2, 12)                 # negative: skipped
        kp[7, :2] = (3.9, 2.2)                 # near the corner: reflect-101 on two sides
        kp[8, :2] = (w - 2.5, 5.7)
        kp[9, :2] = kp[10, :2] = (w // 2, h // 2)          # duplicates (coloured: later one overwrites)
        kp[11, :2] = (w // 2 + 3, h // 2 + 2)              # overlapping blob
        kp[:, 2] = g.uniform(0.2, 1, K)
        out[name] = (h, w, kp)
    h, w = 64, 48
    out["all_skipped"] = (h, w, np.array([[0, 0, 1], [w, h, 1], [-1, -1, 1.0]]))
    out["single"] = (h, w, np.array([[20.5, 33.9, 0.7]]))
    return out


def main():
    if not R.available():
        raise SystemExit("reference tree not found at %s" % R.REF_ROOT)
    ref = reference_rasteriser()
    colors = reference_colors()
    res = {"colors": np.array(colors, np.float64)}
    for name, (h, w, kp) in cases().items():
        K = kp.shape[0]
        res[name + ":kpts"] = kp
        res[name + ":size"] = np.array([h, w])
        # call sites: JointsDataset.py:319-325 (kpts = cond_joints [K, 3] in crop pixels, size = (H, W))
        st = ref.get_stacked_condition(kp.copy(), (h, w), None)               # [H, W, K]
        res[name + ":stacked"] = np.moveaxis(np.asarray(st, np.float64), -1, 0).astype(np.float32)
        pl = ref.get_condition_image(kp.copy(), (h, w))                       # [3, H, W] int
        res[name + ":plain"] = np.asarray(pl).astype(np.int32)
        cols = [colors[i % len(colors)] for i in range(K)]
        co = ref.get_condition_image_colored(kp.copy(), (h, w, 3), colors=cols)   # [H, W, 3]
        res[name + ":colored"] = np.moveaxis(np.asarray(co, np.float64), -1, 0).astype(np.float32)
        print(name, "stacked max %.3f" % res[name + ":stacked"].max(), "plain max %d" % res[name + ":plain"].max(),
              "colored max %.3f" % res[name + ":colored"].max())
    np.savez_compressed(os.path.join(GOLDEN, "condition.npz"), **res)
    print("written", os.path.join(GOLDEN, "condition.npz"))


if __name__ == "__main__":
    main()
