"""TEST INFRASTRUCTURE -- runs the *unmodified* reference modules on the CPU.

Only usable where /root/reference exists (this build container, not the GPU box).  Used by
``oracle/make_golden.py`` to produce the committed fixtures under tests/golden/ and by the
``-m "not gpu"`` tests (when the tree is present) to pin oracle/ against the reference.

What it takes to import the reference here:
  * ``sys.path`` gets ``/root/reference/lib`` (what tools/_init_paths.py:21-24 does);
  * yacs is absent, so the cfg is ``buctd_b200.config.CfgNode`` (dict + attribute access);
  * ``forward`` calls ``x.cuda()`` unconditionally (pose_hrnet_coam.py:495, pose_hrnet.py:453),
    so on a CUDA-less host ``torch.Tensor.cuda`` is patched to identity for the call.
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

REF_ROOT = os.environ.get("BUCTD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


def _import():
    lib = os.path.join(REF_ROOT, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    import importlib
    coam = importlib.import_module("models.pose_hrnet_coam")
    plain = importlib.import_module("models.pose_hrnet")
    inference = importlib.import_module("core.inference")
    loss = importlib.import_module("core.loss")
    transforms = importlib.import_module("utils.transforms")
    return {"pose_hrnet_coam": coam, "pose_hrnet": plain, "inference": inference, "loss": loss,
            "transforms": transforms}


@contextlib.contextmanager
def cpu_cuda_shim():
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build_model(cfg, state_dict=None):
    mods = _import()
    net = mods[cfg.MODEL.NAME].get_pose_net(cfg, is_train=False)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    return net.eval()


def forward(net, x: torch.Tensor) -> torch.Tensor:
    with torch.no_grad(), cpu_cuda_shim():
        return net(x)


def get_final_preds(cfg, heatmaps, center, scale):
    return _import()["inference"].get_final_preds(cfg, heatmaps, center, scale)


def get_max_preds(heatmaps):
    return _import()["inference"].get_max_preds(heatmaps)


def joints_mse_loss(use_target_weight: bool):
    return _import()["loss"].JointsMSELoss(use_target_weight)
