#!/usr/bin/env python
"""Benchmark of the BUCTD conditional top-down hot path (BASELINE.json: crops/sec at 384x288
BUCTD-CoAM-W48, stacked 17-channel condition, batch 64 per GPU, inference forward + decode).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One step = one forward of the network over one batch of synthetic crops + the fused decode.
`value`   : crops/s with the inputs already resident in HBM (whole job, all ranks).
`e2e`     : the same metric through the reference-facing call `model(x_host)` + decode with HOST
            buffers: the pinned-host -> device copy of every step's input and the device -> host
            read of the decoded keypoints are inside the timed region.
`roofline`: the dominant kernel (chain tap-GEMM of the BasicBlock 3x3 convolutions of branch 0): algorithmic
            FLOPs per launch / its launch time measured live with CUDA events, against the MEASURED burst
            bf16/fp16 peak of MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch from the committed
            ncu capture.  `roofline.path_3x3` = the whole path as the fraction of the 3x3-conv tensor-core
            roofline BASELINE.json asks for: crops/s/GPU x 70.34 GFLOP over the measured SUSTAINED peak.
`cpu_baseline`: the oracle port of the reference forward + decode on the host cores (bounded sample).
Only the cpu_baseline / --impl reference legs import oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F3X3_GFLOP = {"coam": 70.34, "prenet": 68.78}      # BASELINE.md section 2 (3x3 convs, per crop)
FTOTAL_GFLOP = {"coam": 88.75, "prenet": 73.17}
METRIC = "crops/sec at 384x288 CoAM-W48"
WORKLOAD = "BUCTD-CoAM-W48 384x288, 17 kpts, stacked condition, batch=%d/GPU, inference forward + decode"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        """started before the warm-up (nvidia-smi needs ~0.5 s to produce its first sample); only samples whose
        arrival time falls inside [mark_begin, mark_end] are reported"""
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        window = "timed region"
        if len(inside) < 2:                      # very short run: fall back to every sample taken under load
            inside = [r for t, r in self.rows if self.t0 is None or t >= self.t0 - 1.0]
            window = "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


CONFIGS = {
    # name: (variant, cond, batch/GPU, passes, workload label)           BASELINE.json configs[2] is the metric's configuration
    "coam64": ("coam", "stacked", 64, 1, "BUCTD-CoAM-W48 384x288, 17 kpts, stacked condition, batch=%d/GPU, inference forward + decode"),
    "prenet32": ("prenet", "colored", 32, 1, "BUCTD-preNet-W48 384x288, 17 kpts, coloured condition, batch=%d/GPU, inference forward + decode"),
    "refine3": ("coam", "stacked", 32, 3, "BUCTD-CoAM-W48 384x288, 17 kpts, 3x iterative self-conditioning refinement (crop fixed), "
                                          "batch=%d/GPU, 3 x (forward + decode + re-rasterise)"),
}


def make_inputs(torch, cfg, batch, device, seed):
    """Synthetic crops on the device: RGB ~ N(0,1); proposals uniform in the crop, ~10 % missing; centers / scales."""
    w, h = cfg.MODEL.IMAGE_SIZE
    k = cfg.MODEL.NUM_JOINTS
    g = torch.Generator(device=device).manual_seed(seed)
    rgb = torch.randn(batch, 3, h, w, generator=g, device=device)
    kp = torch.rand(batch, k, 2, generator=g, device=device) * torch.tensor([w - 16.0, h - 16.0], device=device) + 8.0
    kp[torch.rand(batch, k, generator=g, device=device) < 0.1] = 0.0
    c = torch.tensor([w / 2.0, h / 2.0], device=device) + (torch.rand(batch, 2, generator=g, device=device) * 40 - 20)
    s = torch.tensor([w / 200.0, h / 200.0], device=device) * (0.8 + 0.45 * torch.rand(batch, 1, generator=g, device=device))
    return rgb.contiguous(), kp.contiguous(), c.contiguous(), s.contiguous()


def bind_to_gpu_numa(torch, local):
    """Run this rank (and allocate its pinned host buffers: first touch) on the CPUs of its GPU's NUMA node."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "numa-local cpus %d (%s)" % (len(cpus), path.split("/")[-2])
    except Exception as e:                                   # noqa: BLE001 -- best effort
        return "not bound (%s)" % type(e).__name__
    return "not bound"


def run_native(args):
    import torch
    import torch.distributed as dist
    from buctd_b200 import condition, function, inference, pose_hrnet, pose_hrnet_coam, synth
    from buctd_b200.config import make_cfg
    from buctd_b200.distributed import gather_keypoints

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    variant, cond, B0, passes, label = CONFIGS[args.config]
    B = args.batch or B0
    cfg = make_cfg(variant, 48, (288, 384), cond=cond)
    cfg.TEST.POST_PROCESS = True
    mod = pose_hrnet_coam if variant == "coam" else pose_hrnet
    net = mod.get_pose_net(cfg, is_train=False, dtype=args.dtype)
    net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)     # random-init weights of the architecture
    net = net.cuda().eval()
    K = cfg.MODEL.NUM_JOINTS
    H, W = 384, 288
    # two resident input sets, alternated: 2 x 85 MB of crops > 126 MB L2 (and every step streams ~2.7 GB of workspace)
    sets = [make_inputs(torch, cfg, B, device, 1234 + rank + 1000 * j) for j in range(2)]

    def post(res):
        return gather_keypoints(res, B * world) if world > 1 else res       # the ONE collective: decoded keypoints only

    def step(i):
        rgb, kp, center, scale = sets[i & 1]
        if passes > 1:
            preds, maxvals = function.iterative_refine_from_proposals(cfg, net, rgb, kp, passes, center, scale)
        else:
            preds, maxvals = net.predict_from_proposals(rgb, kp, center, scale, post_process=True)   # head fused into the decode
        return post(torch.cat([preds, maxvals], -1))          # [B,K,3] = 204 B per crop

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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step(i)
        barrier()
        ctx = net._plans[(B, local)]["ctx"]
        l0 = ctx.launch_count()
        sampler.mark_begin()
        ms = timed(lambda n: [step(i) for i in range(n)], args.steps)
        sampler.mark_end()
        launches = ctx.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the public API with HOST buffers ------------------------------------------
        # function.predict_stream(cfg, model, host batches): pinned host crop + proposals -> device (side stream, under the
        # previous batch's forward) -> forward_from_proposals + get_final_preds -> pinned host [B,K,3].
        rgb_dt = torch.float16 if args.e2e_rgb == "f16" else torch.float32
        host = []
        for rgb, kp, c, s in sets:
            host.append((rgb.to(rgb_dt).cpu().pin_memory(), kp.cpu().pin_memory(), c.cpu().pin_memory(), s.cpu().pin_memory()))
        sink = torch.zeros(K, 3)

        def e2e_loop(n):
            if passes > 1:
                out_h = torch.empty(B * world, K, 3).pin_memory()
                for i in range(n):
                    rgb, kp, c, s = host[i & 1]
                    p, m = function.iterative_refine_from_proposals(cfg, net, rgb, kp, passes, c.cuda(non_blocking=True), s.cuda(non_blocking=True))
                    out_h.copy_(post(torch.cat([p, m], -1)), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return
            for res in function.predict_stream(cfg, net, (host[i & 1] for i in range(n)), post=post):
                sink.add_(res[0])                               # the host consumes every result

        e2e_loop(3)
        # three timed passes of exactly K steps each, the fastest one is reported (all are in the line).  (Until the end of
        # round 2 every predict_stream call allocated its device slots on a fresh side stream -- two 85 MB cudaMallocs,
        # ~0.25 s -- which made a 50-step pass read up to 9 ms per step slower; the stream is now cached, profiles/README.md.)
        e2e_passes = [timed(e2e_loop, args.steps) for _ in range(3)]
        ms_e2e = min(e2e_passes)
        h2d = sum(int(t.numel() * t.element_size()) for t in host[0])
        d2h = int(B * world * K * 3 * 4)
        e2e_api = ("function.predict_stream(cfg, model, host batches) -> model.predict_from_proposals"
                   if passes == 1 else "function.iterative_refine_from_proposals(cfg, model, rgb_host, kpts_host, 3, center, scale)")

        # ---- the reference-shaped call: model(x[B, 3+Cc, H, W] float32) with the rasterised planes ------------
        planes = None
        if passes == 1 and world == 1 and not args.quick:
            mode = "stacked" if cond == "stacked" else "colored"
            xs = []
            for rgb, kp, c, s in sets[:1]:
                x = torch.empty(B, 3 + net.arch.in_cond, H, W, device=device)
                x[:, :3] = rgb
                condition.rasterize_condition(kp, H, W, mode, out=x, c_off=3)
                xs.append(x)
            xh = xs[0].cpu().pin_memory()
            out_h = torch.empty(B, K, 3).pin_memory()

            def planes_dev(n):
                for i in range(n):
                    hm = net(xs[0])
                    p, m = inference.decode_device(hm, sets[0][2], sets[0][3], post_process=True)

            def planes_host(n):
                for i in range(n):
                    hm = net(xh.cuda(non_blocking=True))
                    p, m = inference.decode_device(hm, sets[0][2], sets[0][3], post_process=True)
                    out_h.copy_(torch.cat([p, m], -1), non_blocking=True)
                torch.cuda.current_stream().synchronize()

            nn_ = max(5, args.steps // 4)
            planes_dev(2); planes_host(2)
            t_dev, t_host = timed(planes_dev, nn_), timed(planes_host, nn_)
            planes = {"api": "model(x[B,%d,H,W] float32) + decode_device (the reference's forward signature)" % xs[0].shape[1],
                      "value": round(B * nn_ / (t_dev / 1e3), 2), "e2e_value": round(B * nn_ / (t_host / 1e3), 2), "unit": "crops/s",
                      "h2d_bytes_per_step": int(xh.numel() * 4), "steps": nn_,
                      "note": "e2e_value: pinned host tensor -> x.cuda() -> forward -> decode -> host, no overlap"}
            del xs, xh

        kern = time_dominant_conv(torch, net, B, local) if variant == "coam" else None
        pair_kerns = time_pair_kernels(torch, net, B, local) if variant == "coam" and passes == 1 else None
        lib = None
        if world == 1 and passes == 1 and not args.no_gpu_baseline and not args.quick:
            lib = gpu_library_baseline(torch, cfg, net, sets[0], variant, args.lib_batch or B)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    crops = B * world * args.steps
    value = crops / (ms / 1e3)
    per_gpu = value / world
    ach = per_gpu * passes * F3X3_GFLOP[variant] / 1e3
    out = {
        "metric": METRIC if args.config == "coam64" else "crops/sec at 384x288 (%s)" % args.config,
        "value": round(value, 2), "unit": "crops/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": label % B, "name": args.config, "global_batch": B * world, "parallelism": "crops sharded, dp%d" % world,
                   "weights": "random-init (buctd_b200.synth, seed 0)",
                   "api": "model.predict_from_proposals(rgb[B,3,H,W], kpts[B,K,2], center, scale) (forward + fused head/decode), inputs resident in HBM",
                   "l2": "two resident input sets alternate (2 x %d MB > 126 MB L2); each step streams ~GBs of workspace" % (B * 3 * H * W * 4 >> 20),
                   "collective": "one NCCL all_gather of decoded [B,17,3] per step" if world > 1 else "none",
                   "host_binding": numa},
        "roofline": dict(roofline(kern, peaks, ach, per_gpu * passes, variant), **({"cta_pair_kernels": pair_kerns} if pair_kerns else {})),
        "e2e": {"value": round(crops / (ms_e2e / 1e3), 2), "unit": "crops/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4), "api": e2e_api, "rgb_host_dtype": args.e2e_rgb,
                "passes_ms_per_step": [round(v / args.steps, 4) for v in e2e_passes]},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if planes is not None:
        out["planes_api"] = planes
    if lib is not None:
        out["gpu_library_baseline"] = lib["rates"]
        out["parity"] = lib["parity"]
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        out["cpu_baseline"] = cpu_baseline(sample_crops=4, iters=2, variant=variant, cond=cond)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def gpu_library_baseline(torch, cfg, net, inputs, variant, batch):
    """The bar BASELINE.md section 3 names: the reference network on the SAME B200 through PyTorch eager (cuDNN / cuBLAS / ATen)
    -- (fp32) TF32 off, the numerical oracle; (tf32) PyTorch's default for the reference's fp32 model (cuDNN TF32 convolutions);
    (bf16_cl) autocast(bfloat16) + channels_last, the strongest library configuration.  /root/reference does not exist on the GPU
    box, so the network is oracle/hrnet_oracle.forward (the functional torch restatement pinned to the reference by
    tests/test_oracle_vs_golden.py) -- the same ATen/cuDNN calls the reference's nn.Modules make.  Forward only (the reference
    then decodes on the host: cpu_baseline times that).  Also the parity of THIS run: max-abs heat-map error of the CUDA path
    against the fp32 (TF32 off) library forward on the same crops."""
    import contextlib
    from buctd_b200 import condition
    from oracle import hrnet_oracle
    rgb, kp, c, s = inputs
    B = min(batch, rgb.shape[0])
    H, W = rgb.shape[2], rgb.shape[3]
    mode = "stacked" if cfg.DATASET.STACKED_CONDITION else "colored"
    x = torch.empty(B, 3 + net.arch.in_cond, H, W, device=rgb.device)
    x[:, :3] = rgb[:B]
    condition.rasterize_condition(kp[:B], H, W, mode, out=x, c_off=3)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    rates, ref_hm = {}, None
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True                      # cfg.CUDNN.BENCHMARK, tools/test.py:84
    try:
        for name in ("fp32", "tf32", "bf16_cl"):
            torch.backends.cudnn.allow_tf32 = name != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = False     # PyTorch default
            xin, sdn, amp = x, sd, contextlib.nullcontext()
            if name == "bf16_cl":
                xin = x.contiguous(memory_format=torch.channels_last)
                sdn = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
                amp = torch.autocast("cuda", dtype=torch.bfloat16)
            with amp:
                for _ in range(2):
                    hm = hrnet_oracle.forward(sdn, cfg, xin)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 3
                e0.record()
                for _ in range(n):
                    hm = hrnet_oracle.forward(sdn, cfg, xin)
                e1.record()
                torch.cuda.synchronize()
            rates[name] = {"value": round(B * n / (e0.elapsed_time(e1) / 1e3), 2), "unit": "crops/s", "batch": B, "steps": n}
            if name == "fp32":
                ref_hm = hm.float().clone()
            else:
                rates[name]["max_abs_err_vs_fp32"] = float((hm.float() - ref_hm).abs().max())
            del hm
            torch.cuda.empty_cache()
        mine = net.forward_from_proposals(rgb[:B], kp[:B])
        mine2 = net(x)
        parity = {"max_abs_err_vs_fp32_library": float((mine - ref_hm).abs().max()),
                  "max_abs_err_planes_api": float((mine2 - ref_hm).abs().max()),
                  "ref_abs_max": float(ref_hm.abs().max()), "crops": B, "dtype": net.compute_dtype,
                  "argmax_agree": float((mine.flatten(2).argmax(-1) == ref_hm.flatten(2).argmax(-1)).float().mean()),
                  "bound": "1e-2 for the 16-bit path (north_star), heat-maps of O(1)"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    rates["what"] = ("reference network through PyTorch eager on this GPU (oracle/hrnet_oracle.forward: functional restatement, same "
                     "ATen/cuDNN calls), forward only, batch %d" % B)
    return {"rates": rates, "parity": parity}


def dominant_traffic():
    """DRAM bytes per launch of the dominant kernel (the dx-fused chain tap-GEMM of branch 0, 8 convolutions 48->48 at
    batch 64): dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture, summarised by
    tools/ncu_traffic.py into profiles/dominant_traffic.json.  None when the file is missing."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "dominant_traffic.json")
    try:
        with open(path) as f:
            launches = json.load(f)["launches"]
    except (OSError, ValueError, KeyError):
        return None, None
    for o in launches:
        if "tapgemm_chain_kernel<__half, 48" in o["kernel"]:
            return o["dram_read_bytes"] + o["dram_write_bytes"], o["kernel"]
    return None, None


def roofline(kern, peaks, path_tflops, per_gpu, variant="coam"):
    """`roofline` of the bench line: the dominant kernel (the BasicBlock 3x3 convolutions of branch 0: 17.5 % of a
    forward, profiles/README.md) timed alone with CUDA events -> burst peak; plus the whole path against the
    3x3-conv tensor-core roofline BASELINE.json asks for (kernels timed inside a long step -> sustained peak)."""
    out = {"bound": "tensor", "unit": "TFLOP/s", "peak": peaks["tflops_burst"],
           "peak_source": peaks["source"] + " (burst bf16/fp16 dense, MEASURED_PEAKS.json)", "achieved": None, "frac": None,
           "traffic": None}
    if kern is not None:
        traffic, traffic_kernel = dominant_traffic()
        out.update({"achieved": kern["tflops"], "frac": kern["frac_of_burst_peak"], "kernel": kern["name"],
                    "us_per_launch": kern["us_per_launch"], "us_per_conv": kern["us_per_conv"],
                    "algorithmic_flops_per_launch": kern["flops_per_launch"],
                    "traffic": traffic if kern["convs_per_launch"] == 8 and kern["batch"] == 64 else None,
                    "traffic_source": "ncu --set full, profiles/dominant_traffic.json: %s (cold L2, isolated launch)" % traffic_kernel,
                    "hbm_bytes_min_per_launch": kern["hbm_bytes_min_per_conv"] * kern["convs_per_launch"]})
    out["path_3x3"] = {"achieved": round(path_tflops, 2), "peak": peaks["tflops_sustained"],
                       "frac": round(path_tflops / peaks["tflops_sustained"], 4),
                       "what": "forwards/s/GPU x %.2f GFLOP of 3x3 convs per crop (BASELINE.md) over the sustained measured peak" % F3X3_GFLOP[variant],
                       "all_contractions_tflops": round(per_gpu * FTOTAL_GFLOP[variant] / 1e3, 2)}
    return out


def time_pair_kernels(torch, net, B, local):
    """The two CTA-pair (tcgen05 cta_group::2) launch classes, each timed alone with CUDA events like the dominant kernel:
    the channel attention's fc_o of CoAM branch 0 (pairgemm_kernel: Linear(HW, HW) on the planar [B*C, HW] matrix, run on
    its lane's share of the SMs) and the 8-convolution chain of branch 2 (C = 192 @24x18, pair plan of the chain kernel)."""
    st = net._plans[(B, local)]
    plan = st["plan"]
    stream = torch.cuda.current_stream().cuda_stream
    peaks = load_peaks()
    a = net.arch
    out = []
    for label, what, flops in (
            ("stage2_att.att_layers.0.channel_attention_module.pa.fc_o",
             "pairgemm_kernel: Linear(H*W, H*W) of the channel attention, CoAM branch 0 (M = B*48, K = 97*73 padded pixels, N = 96*72)",
             2.0 * B * 48 * (97 * 73) * (96 * 72)),
            ("stage3.0.branches.2:chain",
             "tapgemm_chain_kernel<192, pair>: 8 convolutions 3x3 192->192 @24x18 (branch 2)",
             8 * 2.0 * B * 24 * 18 * 192 * 192 * 9)):
        idx = [i for i, l in enumerate(plan.labels) if l == label]
        if not idx:
            continue
        i = idx[0]
        for _ in range(5):
            plan.run(stream, i, i + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            plan.run(stream, i, i + 1)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        out.append({"kernel": what, "plan_op": label, "us_per_launch": round(us, 2), "algorithmic_flops_per_launch": flops,
                    "tflops": round(flops / us / 1e6, 2), "frac_of_burst_peak": round(flops / us / 1e6 / peaks["tflops_burst"], 4)})
    if out:
        out[0]["note"] = ("runs on the SM share of its plan lane (84 of 148 SMs: the three DA modules run side by side); alone on "
                          "the GPU the same product takes 209 us = 1439 TFLOP/s (tools/pairgemm_bench.py, profiles/README.md)")
    return out or None


def time_dominant_conv(torch, net, B, local):
    """The most expensive launch class, timed alone with CUDA events: the 3x3 s1 48->48 convolutions at 96x72 of the
    BasicBlocks of branch 0 (64 per forward).  With chain launches (default) one op = the 8 convolutions of one
    module's branch; per-launch algorithmic FLOPs = n_convs * 2 * B*96*72 * 48*48*9."""
    st = net._plans[(B, local)]
    plan = st["plan"]
    # (a chain whose live set exceeds the L2 is cut into crop ranges: consecutive ops "...:chain:crops<i>", timed together)
    idx = [i for i, l in enumerate(plan.labels) if l == "stage3.0.branches.0:chain" or l.startswith("stage3.0.branches.0:chain:crops")]
    nconv = 8
    if not idx:
        idx = [i for i, l in enumerate(plan.labels) if l.startswith("stage3.0.branches.0.") and l.endswith("conv1")][:1]
        nconv = 1
    if not idx:
        return None
    i, j = idx[0], idx[-1] + 1
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        plan.run(stream, i, j)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        plan.run(stream, i, j)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flops = nconv * 2.0 * B * 96 * 72 * 48 * 48 * 9
    peaks = load_peaks()
    return {"name": "tap-GEMM 3x3 s1 48->48 @96x72 (BasicBlocks of branch 0), %d convolution(s) per launch%s" % (
                nconv, ", in %d crop-range launches" % (j - i) if j - i > 1 else ""),
            "convs_per_launch": nconv, "batch": B, "flops_per_launch": flops, "us_per_launch": round(us, 2), "us_per_conv": round(us / nconv, 2),
            "tflops": round(flops / us / 1e6, 2), "frac_of_burst_peak": round(flops / us / 1e6 / peaks["tflops_burst"], 4),
            "hbm_bytes_min_per_conv": int(2 * B * 97 * 73 * 48 * 2), "note": "back-to-back launches: activations stay in L2"}


def cpu_baseline(sample_crops=4, iters=2, variant="coam", cond="stacked"):
    """The oracle port of the reference forward + get_final_preds on the host cores (reported baseline only)."""
    import torch
    from buctd_b200 import synth
    from buctd_b200.arch import HRNetArch
    from buctd_b200.config import make_cfg
    from oracle import condition_oracle, decode_oracle, hrnet_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg(variant, 48, (288, 384), cond=cond)
    sd = synth.synth_state_dict(HRNetArch(cfg), 0)
    rgb = synth.synth_rgb(sample_crops, 288, 384)
    kp = synth.synth_keypoints(sample_crops, 17, 288, 384)
    if cond == "stacked":
        cnd = np.stack([condition_oracle.stacked_condition(kp[b], 384, 288) for b in range(sample_crops)]).astype(np.float32)
    else:
        cnd = np.stack([condition_oracle.colored_condition(kp[b], 384, 288, condition_oracle.default_colors()) for b in range(sample_crops)]).astype(np.float32)
    x = torch.from_numpy(np.concatenate([rgb, cnd], 1))
    c, s = synth.synth_center_scale(sample_crops, 288, 384)
    with torch.no_grad():
        hrnet_oracle.forward(sd, cfg, x)                       # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            hm = hrnet_oracle.forward(sd, cfg, x).numpy()
            decode_oracle.get_final_preds(True, hm, c, s)
        dt = (time.perf_counter() - t0) / iters
    return {"value": round(sample_crops / dt, 3), "unit": "crops/s", "cores": cores, "kind": "port",
            "sample": "%d crops x %d iterations of the same workload (fp32 oracle port of the reference forward + decode, "
                      "torch CPU threads=%d)" % (sample_crops, iters, cores)}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; /root/reference does not exist on the
    GPU box and the reference has no native code to compile).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from buctd_b200 import synth
    from buctd_b200.arch import HRNetArch
    from buctd_b200.config import make_cfg
    from oracle import condition_oracle, decode_oracle, hrnet_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.ref_crops
    cfg = make_cfg("coam", 48, (288, 384), cond="stacked")
    sd = synth.synth_state_dict(HRNetArch(cfg), 0)
    rgb = synth.synth_rgb(n, 288, 384)
    kp = synth.synth_keypoints(n, 17, 288, 384)
    cond = np.stack([condition_oracle.stacked_condition(kp[b], 384, 288) for b in range(n)]).astype(np.float32)
    x = torch.from_numpy(np.concatenate([rgb, cond], 1))
    c, s = synth.synth_center_scale(n, 288, 384)

    def step():
        hm = hrnet_oracle.forward(sd, cfg, x).numpy()
        decode_oracle.get_final_preds(True, hm, c, s)

    # bound the whole run to a few minutes whatever K the driver passes
    with torch.no_grad():
        step()
        t0 = time.perf_counter(); step(); per = time.perf_counter() - t0
        steps = max(1, min(args.steps, int(150.0 / max(per, 1e-3))))
        warm = max(0, min(args.warmup, 3) - 2)
        for _ in range(warm):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    v = round(n * steps / dt, 3)
    sample = "%d crops per step x %d timed steps (fp32 oracle port of the reference forward + decode, %d CPU threads)" % (n, steps, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm + 2, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % 64, "sample_crops_per_step": n},
        "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("BUCTD_DTYPE", "fp16"), choices=["fp16", "bf16"])
    ap.add_argument("--config", default="coam64", choices=sorted(CONFIGS) + ["train"],
                    help="coam64 = BASELINE.json's metric configuration (configs[2]); prenet32 / refine3 / train = configs[1] / [3] / [4]")
    ap.add_argument("--batch", type=int, default=0, help="crops per GPU (0: the configuration's own)")
    ap.add_argument("--e2e-rgb", default="f32", choices=["f32", "f16"], help="dtype of the host crop the e2e loop ships")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--lib-batch", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="main + e2e numbers only (no baselines / planes API leg)")
    ap.add_argument("--ref-crops", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.dtype_given = any(a.startswith("--dtype") for a in sys.argv[1:])
    # Exactly ONE line on stdout: libraries that write to file descriptor 1 themselves (NCCL prints its version banner
    # there when a communicator is created) are sent to stderr; the JSON line goes to the original stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "train":
        from buctd_b200 import train_bench
        train_bench.run(args, ClockSampler, load_peaks)
    else:
        run_native(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
