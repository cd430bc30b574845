"""Dense-traffic probe (BASELINE configs[4]): ~10 k overlapping squitters/s at 4 Msps. Parity vs the oracle and timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rate = 4e6
n = 1 << logn
nb = int(10_000 * n / rate)
t = time.time()
sc = synth.make_scene(rate, n, nb, 123, garble_frac=0.2, fruit=nb // 4, snr_db=(4.0, 30.0))
print("scene %d samples, %d bursts (%.1fs)" % (n, nb, time.time() - t))
t = time.time()
want = co.Port().run_iq(sc.iq, rate, 7.0, True, co.MA_SLIDING64)
print("oracle: %d det %d msgs (%.1fs)" % (len(want.index), len(want.msgs), time.time() - t))
iq = torch.from_numpy(sc.iq).cuda()
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True)
rx._ctx.call("amb_enable_timing", 1)
for it in range(3):
    rx.reset(); q.flush(); rx._slicer._first = True
    rx.process(iq, flush=True, collect=False)
    st = rx.stats(); rx.drain()
    print("cuda it%d: scan %.3f ms total %.3f ms (%.1f GS/s) cand %d real %d det %d msgs %d fallback %d | det equal %s payload equal %s" % (
        it, st.ms_scan, st.ms_total, n / st.ms_total / 1e6, st.candidates, st.candidates_real, st.detections, len(q.strings()),
        st.resolver_fallback, [f.sample_index for f in rx.frames] == [int(x) for x in want.index],
        [m.split()[:2] for m in q.strings()] == [m.split()[:2] for m in want.msgs]))
