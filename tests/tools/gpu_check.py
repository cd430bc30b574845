"""Developer diagnostic run on the GPU box: parity of the CUDA path vs the oracle with verbose output."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co

port = co.Port()
ok_all = True

def run_cuda(iq, rate, thr, pmf, chunks=None, resolver=0):
    q = am.msg_queue()
    rx = am.rx_path(rate, thr, q, use_pmf=pmf)
    rx._ctx.call("amb_set_option", b"resolver", resolver)
    frames = []
    if chunks is None:
        rx.process(iq, flush=True); frames += rx.frames
    else:
        pos = 0
        for k, c in enumerate(chunks):
            last = pos + c >= iq.size // 2
            rx.process(iq[2 * pos: 2 * (pos + c)], flush=last); frames += rx.frames
            pos += c
            if last: break
    st = rx.stats()
    return q.strings(), frames, st, rx

for rate, n, nb, pmf in ((4e6, 1 << 20, 60, True), (2e6, 1 << 20, 60, True), (10e6, 1 << 21, 60, True),
                         (20e6, 1 << 22, 60, True), (4e6, 1 << 20, 60, False), (5e6, 1 << 20, 40, True)):
    sc = synth.make_scene(rate, n, nb, seed=int(rate / 1e6) + 7)
    t = time.time()
    o = port.run_iq(sc.iq, rate, 7.0, pmf, co.MA_CANONICAL)
    t_or = time.time() - t
    msgs, frames, st, rx = run_cuda(sc.iq, rate, 7.0, pmf)
    same = msgs == o.msgs
    det_idx = [f.sample_index for f in frames]
    same_det = det_idx == [int(x) for x in o.index]
    ok_all &= same and same_det
    print("rate %.0f pmf %d n %d: oracle %d det %d msgs (%.2fs) | cuda %d det %d msgs | cand %d real %d | msgs_equal %s det_equal %s"
          % (rate, pmf, n, len(o.index), len(o.msgs), t_or, len(frames), len(msgs), st.candidates, st.candidates_real, same, same_det))
    if not same:
        so, sc_ = set(o.msgs), set(msgs)
        print("   missing", list(so - sc_)[:3], "extra", list(sc_ - so)[:3])
        print("   first oracle", o.msgs[:2], "first cuda", msgs[:2])
    if not same_det:
        a, b = set(det_idx), set(int(x) for x in o.index)
        print("   det missing", sorted(b - a)[:5], "det extra", sorted(a - b)[:5])
    msgs1, frames1, st1, _ = run_cuda(sc.iq, rate, 7.0, pmf, resolver=1)
    r_ok = msgs1 == o.msgs
    ok_all &= r_ok
    print("   sequential resolver equal:", r_ok, "fallback(par):", st.resolver_fallback)
    # streaming in ragged chunks
    rng = np.random.default_rng(1)
    chunks = list(rng.integers(1, 200000, 400))
    msgs2, frames2, _, _ = run_cuda(sc.iq, rate, 7.0, pmf, chunks)
    s_ok = msgs2 == o.msgs
    ok_all &= s_ok
    print("   streaming ragged chunks equal:", s_ok, len(msgs2))
    if not s_ok:
        so, sc_ = set(o.msgs), set(msgs2)
        print("   missing", list(so - sc_)[:3], "extra", list(sc_ - so)[:3])

# a > 2^24-sample gap between packets: float rounding at preamble_impl.cc:237 -> sequential fallback
rate = 10e6
sc = synth.make_scene(rate, (1 << 25) + 5_000_001, 0, 3, starts=[1_000_000.4, 1_000_000.4 + (1 << 24) + 4_000_001, 33_000_000.2], amplitude=0.3)
o = port.run_iq(sc.iq, rate, 7.0, True, co.MA_SLIDING64)
msgs, frames, st, rx = run_cuda(sc.iq, rate, 7.0, True)
print("long gap: oracle det", [int(x) for x in o.index], "cuda det", [f.sample_index for f in frames], "fallback", st.resolver_fallback,
      "payload equal", [m.split()[:2] for m in msgs] == [m.split()[:2] for m in o.msgs])
ok_all &= [f.sample_index for f in frames] == [int(x) for x in o.index]

# throughput probe, device-resident input
import torch
for rate, logn in ((4e6, 26), (4e6, 28)):
    n = 1 << logn
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    iq = torch.randn(2 * n, device="cuda", generator=g) * 0.01
    q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True)
    rx._ctx.call("amb_enable_timing", 1)
    for it in range(4):
        rx.reset()
        rx.process(iq, flush=True, collect=False)
        st = rx.stats()
        nm = rx.drain()
        print("n=2^%d it %d: scan %.3f ms (%.1f GS/s, %.1f GB/s) total %.3f ms (%.1f GS/s) cand %d det %d msgs %d"
              % (logn, it, st.ms_scan, n / st.ms_scan / 1e6, 8 * n / st.ms_scan / 1e6, st.ms_total, n / st.ms_total / 1e6,
                 st.candidates, st.detections, nm))
    del iq
print("ALL OK" if ok_all else "MISMATCH")
