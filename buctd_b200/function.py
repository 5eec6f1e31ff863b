"""Device-side mirror of the hot section of the reference's validation loop (``validate``,
/root/reference/lib/core/function.py:200-270): forward, flip test (input flip -> forward -> flip_back ->
SHIFT_HEATMAP -> average) and ``get_final_preds``, with nothing leaving the GPU until the decoded keypoints
(204 B per crop) -- and the synthetic fixed-crop iterative refinement of BASELINE.json configs[3]
(SURVEY.md 8d "Config 4": pass n+1 is conditioned on the rasterised decode of pass n)."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import condition, inference, transforms

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]   # lib/dataset/coco.py:66-67


@torch.no_grad()
def forward_flip_test(config, model, input: torch.Tensor, flip_pairs: Sequence[Sequence[int]] = COCO_FLIP_PAIRS) -> torch.Tensor:
    """``output`` of validate() after the FLIP_TEST block (function.py:207-236); reads TEST.FLIP_TEST,
    TEST.SHIFT_HEATMAP, MODEL.CONDITIONAL_TOPDOWN, DATASET.STACKED_CONDITION."""
    x = input.cuda()
    output = model(x)
    if not config.TEST.FLIP_TEST:
        return output
    if config.MODEL.CONDITIONAL_TOPDOWN:
        if x.shape[1] == 6 and not config.DATASET.STACKED_CONDITION:
            raise NotImplementedError("a 3-channel (coloured) condition is re-rasterised from mirrored proposals in the "
                                      "reference (transforms.py:38-47): build it with condition.rasterize_condition")
        x_f = transforms.flip_input_device(x, flip_pairs, stacked=bool(config.DATASET.STACKED_CONDITION))
    else:
        x_f = transforms.flip_channels_device(x, list(range(x.shape[1])))
    output_flipped = model(x_f)
    return transforms.flip_merge_device(output, output_flipped, flip_pairs, bool(config.TEST.SHIFT_HEATMAP))


@torch.no_grad()
def validate_batch(config, model, input, center, scale, flip_pairs=COCO_FLIP_PAIRS):
    """forward (+ flip test) + get_final_preds on the device: returns (preds [B,K,2], maxvals [B,K,1]) CUDA tensors"""
    output = forward_flip_test(config, model, input, flip_pairs)
    return inference.decode_device(output, center, scale, post_process=bool(config.TEST.POST_PROCESS))


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
