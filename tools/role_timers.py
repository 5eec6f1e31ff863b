#!/usr/bin/env python
"""Debug: build with `make -C buctd_b200/csrc EXTRA=-DBUCTD_ROLE_TIMERS`, then run one slab conv and print
where the MMA / epilogue warps spend their cycles."""
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from buctd_b200 import _lib as L  # noqa: E402
from gpu_utils import TDT, tapgemm, to_act16  # noqa: E402

lib = L.load()
CASES = [(48, 48, 3, 64, 96, 72, False), (48, 48, 3, 64, 96, 72, True), (96, 96, 3, 64, 48, 36, True),
         (192, 192, 3, 64, 24, 18, True), (384, 384, 3, 64, 12, 9, True), (64, 256, 1, 64, 96, 72, True),
         (48, 48, 3, 8, 96, 72, False), (48, 48, 3, 16, 96, 72, False), (48, 48, 3, 32, 96, 72, False), (48, 48, 3, 128, 96, 72, False),
         (96, 96, 3, 16, 48, 36, False), (96, 96, 3, 128, 48, 36, False)]
for (cin, cout, k, B, h, w, res) in CASES:
    dt = L.F16
    x = torch.randn(B, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(k * k * cin)
    bias = torch.randn(cout, device="cuda")
    a16 = to_act16(x, dt)
    out = torch.zeros(B, h + 1, w + 1, cout, dtype=TDT[dt], device="cuda")
    r16 = to_act16(torch.randn(B, cout, h, w, device="cuda"), dt) if res else None
    for _ in range(3):
        tapgemm(dt, a16, wt, bias, out, H=h, W=w, relu=True, residual=r16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tapgemm(dt, a16, wt, bias, out, H=h, W=w, relu=True, residual=r16)
    e1.record()
    torch.cuda.synchronize()
    print("%d->%d k%d B%d %dx%d res=%d: %.1f us (one launch incl. host prepare)" % (cin, cout, k, B, h, w, res, e0.elapsed_time(e1) * 1e3))
    buf = (ctypes.c_ulonglong * (148 * 16))()
    if hasattr(lib, "buctd_debug_role_cycles"):
        lib.buctd_debug_role_cycles(buf, 148 * 16)
        a = np.array(buf[:], dtype=np.float64).reshape(148, 16).mean(0)
        rows = B * (h + 1) * (w + 1)
        nmma = -(-rows // 128) * k * k * (-(-cin // 16)) * (-(-cout // 192) if cout > 256 else 1) / 148.0
        print("%d->%d k%d B%d %dx%d res=%d | MMA warp total %.0f cyc (%.0f per MMA, %.0f MMAs/CTA): wait acc_empty %.0f, slab %.0f, W %.0f, issue %.0f, commit %.0f | epilogue total %.0f, wait acc_full %.0f"
              % (cin, cout, k, B, h, w, res, a[0], a[0] / nmma, nmma, a[1], a[2], a[3], a[6], a[7], a[4], a[5]))
        print("      epilogue thread64: (unused %.0f) stage-ready wait %.0f, compute %.0f (unused %.0f %.0f) | thread101: compute %.0f (unused %.0f)"
              % (a[8], a[9], a[10], a[11], a[12], a[13], a[14]))
    else:
        print("%d->%d: %.1f us (library built without BUCTD_ROLE_TIMERS)" % (cin, cout, e0.elapsed_time(e1) * 1e3))
