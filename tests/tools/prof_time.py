"""Timing probe (device-resident): scan-kernel and whole-call time, plus a parity check vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 4e6
sc = synth.make_scene(rate, 1 << 20, 60, seed=11)
want = co.Port().run_iq(sc.iq, rate, 7.0, True, co.MA_CANONICAL).msgs
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True); rx.process(sc.iq, flush=True)
ok = q.strings() == want
n = 1 << logn
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randn(2 * n, device="cuda", generator=g) * 0.01
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True)
rx._ctx.call("amb_enable_timing", 1)
sc_ms, tot = [], []
for it in range(6):
    rx.reset(); rx.process(iq, flush=True, collect=False)
    st = rx.stats(); rx.drain()
    if it >= 2: sc_ms.append(st.ms_scan); tot.append(st.ms_total)
# back-to-back steps on torch's stream (what bench.py times)
rx._ctx.use_stream(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    rx.reset(); rx.process(iq, flush=True, collect=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for it in range(K):
    rx.reset(); rx.process(iq, flush=True, collect=False)
rx._ctx.join(); e1.record(); torch.cuda.synchronize()
nm = rx.drain()
ov = rx._ctx.scan_times_ms(K)
b2b = e0.elapsed_time(e1) / K
print("parity %s scan %.3f ms (%.1f GS/s, %.0f GB/s) single-call %.3f ms  back-to-back %.3f ms/step (%.1f GS/s; scan under overlap %.3f ms) msgs %d" % (
    ok, np.mean(sc_ms), n / np.mean(sc_ms) / 1e6, 8 * n / np.mean(sc_ms) / 1e6, np.mean(tot), b2b, n / b2b / 1e6, np.mean(ov), nm))
