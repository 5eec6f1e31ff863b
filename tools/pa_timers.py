#!/usr/bin/env python
"""Debug (library built with EXTRA=-DBUCTD_ROLE_TIMERS): role timers of the tcgen05 position attention."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from buctd_b200 import _lib as L
from gpu_utils import TDT, ctx, stream, to_act16
lib = L.load()
B, C, h, w = 64, 48, 96, 72
dt = L.F16
for scale in (1.0, 6.0):
    q = to_act16(torch.randn(B, C, h, w, device="cuda") * scale, dt)
    kv = to_act16(torch.randn(B, 2 * C, h, w, device="cuda"), dt)
    o = torch.zeros(B, h + 1, w + 1, C, dtype=TDT[dt], device="cuda")
    a = L.PosAttArgs()
    a.dtype, a.batch, a.h, a.w, a.c = dt, B, h, w, C
    a.q_cs, a.k_cs, a.v_cs, a.o_cs = C, 2 * C, 2 * C, C
    a.q, a.k, a.v, a.o = q.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * C, o.data_ptr()
    for _ in range(2):
        ctx().run(L.OP_POS_ATT, a, stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx().run(L.OP_POS_ATT, a, stream()); e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    if not hasattr(lib, "buctd_debug_pa_cycles"):            # product build: time only
        print("q scale %.1f: %.1f us" % (scale, e0.elapsed_time(e1) * 1e3))
        continue
    lib.buctd_debug_pa_cycles(buf)
    v = list(buf)
    nt = max(v[11], 1)
    print("q scale %.1f: %.1f us | per tile: MMA thread total %d (wait K %d, wait P %d, wait V %d) | softmax total %d (bar %d, wait S %d, wait O %d, main pass %d [tmem ld wait %d, ld issue %d, chunks %d], wait V %d), redo %d of %d tiles"
          % (scale, e0.elapsed_time(e1) * 1e3, v[0] // nt, v[1] // nt, v[2] // nt, v[3] // nt, v[4] // nt, v[5] // nt, v[6] // nt, v[7] // nt, v[8] // nt, v[12] // nt, v[13] // nt, v[14] // nt, v[9] // nt, v[10], nt))
