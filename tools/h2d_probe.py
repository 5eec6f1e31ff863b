import torch, time
x = torch.randn(64, 3, 384, 288).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for rep in range(6):
    ts = []
    for i in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s); d.copy_(x, non_blocking=True); e1.record(s)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("pass %d: 85 MB H2D ms min %.2f median %.2f max %.2f -> %.1f GB/s median" % (rep, ts[0], ts[20], ts[-1], x.numel()*4/ts[20]/1e6), flush=True)
    time.sleep(0.5)
# under a concurrent compute load on the default stream
a = torch.randn(8192, 8192, device="cuda", dtype=torch.float16)
for rep in range(4):
    ts = []
    for i in range(20):
        for _ in range(6): a @ a
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s); d.copy_(x, non_blocking=True); e1.record(s)
        e1.synchronize(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("under matmul load pass %d: min %.2f median %.2f max %.2f ms" % (rep, ts[0], ts[10], ts[-1]), flush=True)
