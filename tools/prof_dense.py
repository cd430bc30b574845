"""Dense-traffic workload (BASELINE configs[4]) for ncu and for plain timing: device-resident scene, a few passes.
    python tools/prof_dense.py LOG2N [RATE] [ITERS] [time]
With `time`: per-call stage times (library events) and back-to-back step time; without: just the passes (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 4e6
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
timing = len(sys.argv) > 4 and sys.argv[4] == "time"
n = 1 << logn
nb = int(10_000 * n / rate)
iq, _, _ = synth.make_scene_device(rate, n, nb, 5, torch.device("cuda"), garble_frac=0.2, fruit=nb // 4, snr_db=(4.0, 30.0))
torch.cuda.synchronize()
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True)
if timing:
    rx._ctx.call("amb_enable_timing", 1)
for it in range(iters):
    rx.reset()
    rx.process(iq, flush=True, collect=False)
    st = rx.stats(); nm = rx.drain(); q.flush()
    if timing:
        print("it%d scan %.3f ms total %.3f ms cand %d real %d det %d msgs %d fallback %d" % (
            it, st.ms_scan, st.ms_total, st.candidates, st.candidates_real, st.detections, nm, st.resolver_fallback))
if timing:
    rx._ctx.use_stream(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for it in range(K):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    rx._ctx.join(); e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / K
    sc = float(np.mean(rx._ctx.scan_times_ms(K)))
    rx.drain()
    print("dense 2^%d @ %.0f Msps: back-to-back %.3f ms/step = %.1f GS/s = %.0f GB/s; scan under overlap %.3f ms" % (
        logn, rate / 1e6, b2b, n / b2b / 1e6, 8 * n / b2b / 1e6, sc))
print("done")
