"""bench.py --config train: BASELINE.json configs[4] -- BUCTD-CoAM-W48 384x288 bf16 training step (JointsMSELoss +
backward + Adam), 32 crops per GPU, data parallel with ONE NCCL all-reduce (mean) of the flat gradient buffer per step
(the reference: torch.nn.DataParallel + torch.optim.Adam, tools/train.py:147, lib/core/function.py:113-138).

One step = TrainStep.step(): H2D-free (inputs resident), training-mode forward, loss, backward, all-reduce, Adam, weight
re-pack.  `value` = crops/s over all ranks (max-over-ranks CUDA-event time); `e2e` = the same with the batch (crop planes,
targets, weights) coming from pinned host memory every step and the loss read back."""
from __future__ import annotations

import json
import os

import numpy as np

TRAIN_GFLOP = 3 * 88.75          # ~3x the forward's 88.75 GFLOP per crop


def run(args, ClockSampler, load_peaks):
    import torch
    import torch.distributed as dist
    from buctd_b200 import condition, pose_hrnet_coam, synth
    from buctd_b200.config import make_cfg
    from buctd_b200.train import TrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    B = args.batch or 32
    dtype = args.dtype if args.dtype_given else "bf16"
    cfg = make_cfg("coam", 48, (288, 384), cond="stacked")
    net = pose_hrnet_coam.get_pose_net(cfg, is_train=True, dtype=dtype)
    net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)
    net = net.cuda().train()
    ts = TrainStep(net, B, lr=1e-3, dtype=dtype, loss_scale=1024.0 if dtype == "fp16" else 1.0, drop_p=0.1, seed=1 + rank)
    H, W, K = 384, 288, 17
    g = torch.Generator(device=device).manual_seed(777 + rank)
    sets = []
    for j in range(2):
        x = torch.empty(B, 3 + K, H, W, device=device)
        x[:, :3] = torch.randn(B, 3, H, W, generator=g, device=device)
        kp = torch.rand(B, K, 2, generator=g, device=device) * torch.tensor([W - 16.0, H - 16.0], device=device) + 8.0
        condition.rasterize_condition(kp, H, W, "stacked", out=x, c_off=3)
        # targets: the reference's sigma-3 Gaussian stamps around (jittered) keypoints, 20 % of the joints unlabelled
        yy, xx = torch.meshgrid(torch.arange(H // 4, device=device), torch.arange(W // 4, device=device), indexing="ij")
        c = (kp / 4.0 + torch.randn(B, K, 2, generator=g, device=device)).view(B, K, 2, 1, 1)
        tgt = torch.exp(-((xx - c[:, :, 0]) ** 2 + (yy - c[:, :, 1]) ** 2) / (2 * 3.0 ** 2)).contiguous()
        tw = (torch.rand(B, K, 1, generator=g, device=device) < 0.8).float()
        sets.append((x, tgt, tw))

    ar_ms = [0.0]
    if world > 1:
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

        # ONE logical all-reduce of the flat gradient buffer, issued as buckets: the range of the buffer the backward finishes
        # first (the late stages) is reduced asynchronously under the second half of the backward, the rest after it.
        # ev[0] .. ev[1] = what the step still waits for (the exposed part).  BUCTD_TRAIN_BUCKETS=0: one blocking call.
        bucketed = os.environ.get("BUCTD_TRAIN_BUCKETS", "1") != "0"
        lo, hi = ts.set_bucket_split(0.5) if bucketed else (0, 0)
        work = [None]

        def bucket(lo_, hi_):
            work[0] = dist.all_reduce(ts.opt.flat_grad[lo_:hi_], op=dist.ReduceOp.AVG, async_op=True)

        def allreduce():
            ev[0].record()
            if bucketed:
                if lo > 0:
                    dist.all_reduce(ts.opt.flat_grad[:lo], op=dist.ReduceOp.AVG)
                if hi < ts.opt.flat_grad.numel():
                    dist.all_reduce(ts.opt.flat_grad[hi:], op=dist.ReduceOp.AVG)
                work[0].wait()
            else:
                dist.all_reduce(ts.opt.flat_grad, op=dist.ReduceOp.AVG)
            ev[1].record()
        if bucketed:
            ts.on_bucket_ready = bucket
        ts.on_grads_ready = allreduce

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fn(n)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    losses = []

    def loop(n):
        for i in range(n):
            losses.append(ts.step(*sets[i & 1]).clone())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    loop(max(args.warmup, 3))
    barrier()
    l0 = ts.ctx.launch_count()
    sampler.mark_begin()
    ms = timed(loop, args.steps)
    sampler.mark_end()
    launches = ts.ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        torch.cuda.synchronize()
        ar_ms[0] = ev[0].elapsed_time(ev[1])
    first, last = float(losses[0].item()), float(losses[-1].item())

    host = [tuple(t.cpu().pin_memory() for t in s) for s in sets]
    sink = [0.0]

    def e2e_loop(n):
        for i in range(n):
            sink[0] += float(ts.step(*host[i & 1]).item())          # H2D of the batch + D2H of the loss every step

    e2e_loop(2)
    ms_e2e = timed(e2e_loop, args.steps)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    lib = None
    if world == 1 and not args.no_gpu_baseline and not args.quick:
        lib = library_baseline(torch, cfg, net, sets[0], args.lib_batch or B)
    if rank == 0:
        peaks = load_peaks()
        crops = B * world * args.steps
        value = crops / (ms / 1e3)
        per_gpu = value / world
        out = {
            "metric": "crops/sec, training step at 384x288 CoAM-W48 (forward + JointsMSELoss + backward + Adam)",
            "value": round(value, 2), "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic",
            "config": {"workload": "BUCTD-CoAM-W48 384x288 training step, batch=%d/GPU, JointsMSELoss(use_target_weight) + backward + Adam lr 1e-3, "
                                   "attention dropout 0.1, train-mode BatchNorm (per-GPU statistics)" % B,
                       "name": "train", "global_batch": B * world, "parallelism": "data parallel dp%d, one NCCL all-reduce (AVG) of the flat "
                       "%.1f M-element float32 gradient buffer per step%s" % (
                           world, ts.opt.flat_grad.numel() / 1e6,
                           ", issued as buckets: the late stages' range of the buffer (%.1f M elements, final after the first half of the backward) "
                           "reduced under the second half" % ((ts._split[2] - ts._split[1]) / 1e6) if ts._split else ""),
                       "weights": "random-init (buctd_b200.synth, seed 0)", "l2": "two resident batches alternate (2 x %d MB)" % (B * 20 * H * W * 4 >> 20),
                       "loss_first_last": [round(first, 6), round(last, 6)], "workspace_gb": round(ts.ws.numel() / 2 ** 30, 2)},
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": peaks["tflops_sustained"],
                         "achieved": round(per_gpu * TRAIN_GFLOP / 1e3, 2), "frac": round(per_gpu * TRAIN_GFLOP / 1e3 / peaks["tflops_sustained"], 4),
                         "traffic": None, "what": "whole step: crops/s/GPU x 3 x 88.75 GFLOP (SURVEY.md 8a-T) over the sustained measured peak"},
            "e2e": {"value": round(crops / (ms_e2e / 1e3), 2), "unit": "crops/s",
                    "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host[0])), "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / args.steps, 3), "api": "TrainStep.step(input_host, target_host, target_weight_host).item()"},
            "allreduce_ms": round(ar_ms[0], 3) if world > 1 else None,
            "allreduce_note": ("exposed time: from the end of the backward to the end of the reduction (the early bucket overlaps the backward)"
                               if world > 1 and ts._split else None),
            "gpu_launches": int(args.steps * ts.launches_per_step()), "launches_per_step": ts.launches_per_step(),
            "launch_note": "C-ABI launches recorded per step (replayed as one CUDA graph + Adam + re-pack); %d immediate launches counted by the handle" % launches,
            "clocks": clocks,
        }
        if lib is not None:
            out["gpu_library_baseline"] = lib
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def library_baseline(torch, cfg, net, batch, B):
    """The same training step through PyTorch eager on this GPU (cuDNN / cuBLAS / ATen autograd + torch.optim.Adam): the
    reference's own training loop (lib/core/function.py:113-138) with its network restated functionally
    (oracle/train_oracle.py, pinned to the reference by tests/test_train_oracle.py; /root/reference does not exist on the GPU
    box).  (tf32) PyTorch's default for the reference's fp32 model, (bf16) autocast(bfloat16).  Reported baseline only."""
    import contextlib
    import torch.nn.functional as F
    from oracle import hrnet_oracle as O
    from oracle import train_oracle as T
    x, tgt, tw = [t[:B] for t in batch]
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    params = {k: v.float().requires_grad_(v.dtype.is_floating_point and "running_" not in k) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3)
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name in ("tf32", "bf16"):
            torch.backends.cudnn.allow_tf32 = True
            amp = torch.autocast("cuda", dtype=torch.bfloat16) if name == "bf16" else contextlib.nullcontext()

            def step():
                O._CTX.update({"train": True, "running": None, "drop": None, "drop_p": 0.0})
                try:
                    with amp:
                        hm = O.forward(params, cfg, x)
                        hm = F.dropout(hm, 0.0)
                    loss = T.joints_mse(hm.float(), tgt, tw, True)
                finally:
                    O._CTX.update({"train": False})
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                return loss
            try:
                step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 2
                e0.record()
                for _ in range(n):
                    step()
                e1.record()
                torch.cuda.synchronize()
                out[name] = {"value": round(B * n / (e0.elapsed_time(e1) / 1e3), 2), "unit": "crops/s", "batch": B, "steps": n,
                             "ms_per_step": round(e0.elapsed_time(e1) / n, 2)}
            except torch.cuda.OutOfMemoryError:
                out[name] = {"value": None, "note": "out of memory at batch %d" % B}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
    out["what"] = ("the reference's training step through PyTorch eager on this GPU (functional restatement of the network + "
                   "torch.autograd + torch.optim.Adam; attention dropout omitted), batch %d" % B)
    return out
