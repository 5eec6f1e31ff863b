"""yacs-free config node for the BUCTD hot path.

The reference resolves everything through a global yacs ``CfgNode`` that model code reads
both as ``cfg['MODEL']['EXTRA']`` and ``cfg.MODEL.ATT_MODULES``
(/root/reference/lib/models/pose_hrnet_coam.py:281,346) and that is filled from
``experiments/*.yaml`` plus a ``KEY VAL`` override list
(/root/reference/lib/config/default.py:180-207).  yacs is not a dependency here: ``CfgNode``
below is a dict with attribute access that loads the same YAML files and the same override
list, and ``get_cfg_defaults`` carries the defaults of the keys the hot path consumes
(SURVEY.md section 5, row "Config").  Keys outside that set are accepted and kept verbatim so
a reference YAML loads unmodified.
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Iterable, Sequence

import yaml


class CfgNode(dict):
    """dict with attribute access, nested; supports both ``cfg.A.B`` and ``cfg['A']['B']``."""

    def __init__(self, init: dict | None = None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v: Any) -> Any:
        if isinstance(v, dict) and not isinstance(v, CfgNode):
            return cls(v)
        return v

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = self._wrap(value)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    # yacs API subset -------------------------------------------------------------------
    def defrost(self) -> None:  # the reference freezes/defrosts; mutability is not tracked here
        pass

    def freeze(self) -> None:
        pass

    def merge_from_dict(self, other: dict) -> None:
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge_from_dict(v)
            else:
                self[k] = self._wrap(copy.deepcopy(v))

    def merge_from_file(self, path: str) -> None:
        with open(path, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_dict(_tuplify(loaded))

    def merge_from_list(self, opts: Sequence[Any] | None) -> None:
        opts = list(opts or [])
        if len(opts) % 2:
            raise ValueError("override list must be KEY VAL pairs, got %r" % (opts,))
        for key, raw in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    node[p] = CfgNode()
                node = node[p]
            node[parts[-1]] = self._wrap(_decode(raw))


def _tuplify(v: Any) -> Any:
    # yaml writes GPUS as the string "(0,1,2,3)"
    if isinstance(v, dict):
        return {k: _tuplify(x) for k, x in v.items()}
    if isinstance(v, str) and v.startswith("(") and v.endswith(")"):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _decode(raw: Any) -> Any:
    """CLI values arrive as strings: ``True``, ``[False,True,False,False]``, ``1e-3``, paths."""
    if not isinstance(raw, str):
        return raw
    try:
        return ast.literal_eval(raw)
    except (ValueError, SyntaxError):
        return raw


def get_cfg_defaults() -> CfgNode:
    """Defaults for the keys the hot path reads (values as in the reference's default.py)."""
    return CfgNode({
        "OUTPUT_DIR": "", "LOG_DIR": "", "DATA_DIR": "", "GPUS": (0,), "WORKERS": 4,
        "PRINT_FREQ": 20, "AUTO_RESUME": False, "PIN_MEMORY": True, "RANK": 0,
        "CUDNN": {"BENCHMARK": True, "DETERMINISTIC": False, "ENABLED": True},
        "MODEL": {
            "NAME": "pose_hrnet", "INIT_WEIGHTS": True, "PRETRAINED": "", "NUM_JOINTS": 17,
            "TAG_PER_JOINT": True, "TARGET_TYPE": "gaussian",
            "IMAGE_SIZE": [256, 256], "HEATMAP_SIZE": [64, 64], "SIGMA": 2,
            "EXTRA": {"USE_PRE_NET": False, "USE_ATTENTION": False, "FINAL_CONV_KERNEL": 1,
                      "PRETRAINED_LAYERS": ["*"]},
            "ATT_MODULES": [False, False, True, True], "ATT_CHANNEL_ONLY": False,
            "ATTENTION_HEADS": 1, "SELFATT_MODULES": [False, False, False, False],
            "CONDITIONAL_TOPDOWN": False,
        },
        "LOSS": {"USE_OHKM": False, "TOPK": 8, "USE_TARGET_WEIGHT": True,
                 "USE_DIFFERENT_JOINTS_WEIGHT": False},
        "DATASET": {"DATASET": "coco", "ROOT": "", "COLOR_RGB": False, "COLORED": False,
                    "STACKED_CONDITION": False, "FLIP": True, "BU_BBOX_MARGIN": 25},
        "TRAIN": {"LR": 0.001, "LR_FACTOR": 0.1, "LR_STEP": [90, 110], "OPTIMIZER": "adam",
                  "MOMENTUM": 0.9, "WD": 0.0001, "NESTEROV": False, "GAMMA1": 0.99,
                  "GAMMA2": 0.0, "BEGIN_EPOCH": 0, "END_EPOCH": 140, "BATCH_SIZE_PER_GPU": 32,
                  "SHUFFLE": True},
        "TEST": {"BATCH_SIZE_PER_GPU": 32, "FLIP_TEST": False, "POST_PROCESS": False,
                 "SHIFT_HEATMAP": False, "MODEL_FILE": "", "SCALE_THRE": 1.25,
                 "IN_VIS_THRE": 0.0},
        "DEBUG": {"DEBUG": False},
    })


class _Args:
    def __init__(self, cfg: str, opts: Iterable[Any] = (), modelDir: str = "", logDir: str = "",
                 dataDir: str = ""):
        self.cfg, self.opts = cfg, list(opts)
        self.modelDir, self.logDir, self.dataDir = modelDir, logDir, dataDir


def update_config(cfg: CfgNode, args: Any) -> None:
    """Same contract as the reference's ``update_config(cfg, args)``: merge ``args.cfg`` (yaml)
    then ``args.opts`` (KEY VAL list), then join the data-dir relative paths."""
    cfg.merge_from_file(args.cfg)
    cfg.merge_from_list(getattr(args, "opts", None))
    if getattr(args, "modelDir", ""):
        cfg.OUTPUT_DIR = args.modelDir
    if getattr(args, "logDir", ""):
        cfg.LOG_DIR = args.logDir
    if getattr(args, "dataDir", ""):
        cfg.DATA_DIR = args.dataDir
    cfg.DATASET.ROOT = os.path.join(cfg.DATA_DIR, cfg.DATASET.ROOT)
    cfg.MODEL.PRETRAINED = os.path.join(cfg.DATA_DIR, cfg.MODEL.PRETRAINED)
    if cfg.TEST.MODEL_FILE:
        cfg.TEST.MODEL_FILE = os.path.join(cfg.DATA_DIR, cfg.TEST.MODEL_FILE)


def load_config(yaml_path: str | None = None, opts: Sequence[Any] = ()) -> CfgNode:
    cfg = get_cfg_defaults()
    if yaml_path:
        update_config(cfg, _Args(yaml_path, opts))
    else:
        cfg.merge_from_list(opts)
    return cfg


# ---------------------------------------------------------------------------------------
# Built-in HRNet stage tables so BASELINE.json's configurations can be named without a yaml
# file on disk (the GPU box has no /root/reference).  Values restate
# experiments/coco/hrnet/w{32,48}_384x288_adam_lr1e-3.yaml:43-96.
# ---------------------------------------------------------------------------------------
def _stages(width: int) -> dict:
    c = [width, width * 2, width * 4, width * 8]
    st = {}
    for n, (mods, br) in zip((2, 3, 4), ((1, 2), (4, 3), (3, 4))):
        st["STAGE%d" % n] = {"NUM_MODULES": mods, "NUM_BRANCHES": br, "BLOCK": "BASIC",
                             "NUM_BLOCKS": [4] * br, "NUM_CHANNELS": c[:br],
                             "FUSE_METHOD": "SUM"}
    return st


def make_cfg(variant: str = "coam", width: int = 48, image_size=(288, 384), num_joints: int = 17,
             cond: str = "stacked", post_process: bool = True, att_modules=(False, True, False, False)) -> CfgNode:
    """variant: 'plain' | 'prenet' | 'coam'; cond: 'stacked' | 'colored' | 'plain' (1-ch);
    att_modules: MODEL.ATT_MODULES (the shipped scripts use [F,T,F,F]; lib/config/default.py:48 defaults to [F,F,T,T])."""
    cfg = get_cfg_defaults()
    w, h = int(image_size[0]), int(image_size[1])
    extra = {"USE_PRE_NET": variant == "prenet", "USE_ATTENTION": variant == "coam",
             "FINAL_CONV_KERNEL": 1, "PRETRAINED_LAYERS": ["*"]}
    extra.update(_stages(width))
    cfg.merge_from_dict({
        "MODEL": {"NAME": "pose_hrnet_coam" if variant == "coam" else "pose_hrnet",
                  "NUM_JOINTS": num_joints, "IMAGE_SIZE": [w, h], "HEATMAP_SIZE": [w // 4, h // 4],
                  "SIGMA": 3, "EXTRA": extra, "ATT_MODULES": [bool(v) for v in att_modules],
                  "CONDITIONAL_TOPDOWN": variant != "plain"},
        "DATASET": {"STACKED_CONDITION": cond == "stacked" and variant != "plain",
                    "COLORED": cond == "colored" and variant != "plain"},
        "TEST": {"POST_PROCESS": post_process},
    })
    return cfg
