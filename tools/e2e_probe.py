#!/usr/bin/env python
"""Where does a slow end-to-end pass lose its time?  The loop of function.predict_stream (host batch -> side-stream copy ->
predict_from_proposals -> pinned result) with CUDA events around the copy and the forward of every step, several passes.
    python tools/e2e_probe.py [--passes 6] [--steps 40]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from buctd_b200 import pose_hrnet_coam, synth
from buctd_b200.config import make_cfg
from bench import make_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=6)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--reuse-stream", action="store_true")
args = ap.parse_args()
cfg = make_cfg("coam", 48, (288, 384), cond="stacked")
cfg.TEST.POST_PROCESS = True
net = pose_hrnet_coam.get_pose_net(cfg, is_train=False, dtype="fp16")
net.load_state_dict(synth.synth_state_dict(net.arch, 0), strict=True)
net = net.cuda().eval()
dev = torch.device("cuda", 0)
B = 64
sets = [make_inputs(torch, cfg, B, dev, 1234 + 1000 * j) for j in range(2)]
host = [tuple(t.cpu().pin_memory() for t in s) for s in sets]
main = torch.cuda.current_stream(dev)
fixed_side = torch.cuda.Stream(dev)
with torch.no_grad():
    for _ in range(3):
        net.predict_from_proposals(*sets[0], post_process=True)
    torch.cuda.synchronize()
    for p in range(args.passes):
        side = fixed_side if args.reuse_stream else torch.cuda.Stream(dev)
        slots = [None, None]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        outs = [torch.empty(B, 17, 3).pin_memory() for _ in range(2)]
        ev = []

        def stage(i):
            b = i & 1
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(freed[b])
                if slots[b] is None:
                    slots[b] = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0])
                c0.record(side)
                for d, h in zip(slots[b], host[b]):
                    d.copy_(h, non_blocking=True)
                c1.record(side)
                ready[b].record(side)
            return c0, c1

        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        cev = {0: stage(0)}
        pending = None
        for i in range(args.steps):
            b = i & 1
            main.wait_event(ready[b])
            if i + 1 < args.steps:
                cev[i + 1] = stage(i + 1)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(main)
            preds, maxvals = net.predict_from_proposals(*slots[b], post_process=True)
            res = torch.cat([preds, maxvals], 2)
            f1.record(main)
            freed[b].record(main)
            outs[b].copy_(res, non_blocking=True)
            done[b].record(main)
            ev.append((cev[i][0], cev[i][1], f0, f1))
            if pending is not None:
                done[pending].synchronize()
            pending = b
        t1.record()
        torch.cuda.synchronize()
        tot = t0.elapsed_time(t1) / args.steps
        cp = sorted(e[0].elapsed_time(e[1]) for e in ev)
        fw = sorted(e[2].elapsed_time(e[3]) for e in ev)
        gap = sorted(ev[i][3].elapsed_time(ev[i + 1][2]) for i in range(len(ev) - 1))      # end of forward i -> start of forward i+1
        late = sorted(ev[i + 1][1].elapsed_time(ev[i][3]) for i in range(len(ev) - 1))      # copy i+1 done -> forward i done (> 0: in time)
        m = len(ev) // 2
        print("pass %d: %.2f ms/step | copy median %.2f max %.2f | forward median %.2f max %.2f | gap between forwards median %.2f max %.2f | "
              "copy(i+1) finished %.2f ms before forward(i) ended (median; min %.2f)" % (p, tot, cp[m], cp[-1], fw[m], fw[-1], gap[m - 1], gap[-1],
                                                                                      late[m - 1], late[0]), flush=True)
