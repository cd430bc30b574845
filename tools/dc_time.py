import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import gr_air_modes_b200 as am
n = 1 << 26
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randn(2 * n, device="cuda", generator=g) * 0.01 + 0.02
for rate in (4e6, 10e6):
    for dc in (False, True):
        q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True, use_dcblock=dc)
        rx._ctx.use_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            rx.reset(); rx.process(iq, flush=True, collect=False)
        rx._ctx.join(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rx.reset(); rx.process(iq, flush=True, collect=False)
        rx._ctx.join(); e1.record(); torch.cuda.synchronize()
        print("rate %g dcblock %s: %.3f ms per 2^26 samples" % (rate, dc, e0.elapsed_time(e1) / 10))
        rx.close()
