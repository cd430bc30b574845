"""Randomised parity stress on the GPU: rates, thresholds, PMF, chunkings, amplitude scales, silence, NaN/Inf."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co
port = co.Port()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncase = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = 0
t0 = time.time()
for case in range(ncase):
    rate = float(rng.choice([2e6, 2.4e6, 3e6, 4e6, 5e6, 6e6, 8e6, 10e6, 12e6, 16e6, 20e6, 4e6, 4e6, 2e6]))
    pmf = bool(rng.random() < 0.8)
    thr = float(rng.choice([3.0, 5.0, 7.0, 7.0, 9.0, 12.0]))
    n = int(rng.integers(5_000, 400_000))
    nb = int(rng.integers(0, 60))
    sigma = float(rng.choice([0.0, 1e-3, 0.01, 0.01, 0.05]))
    kind = rng.choice(["plain", "plain", "scale_small", "scale_big", "silence", "naninf", "dc", "dense"])
    sc = synth.make_scene(rate, n, nb if kind != "dense" else nb * 10, int(rng.integers(1 << 30)), noise_sigma=sigma,
                          snr_db=(3.0, 40.0), garble_frac=0.2, fruit=int(rng.integers(0, 40)),
                          amplitude=None if sigma > 0 else 0.3)
    iq = sc.iq.copy()
    if kind == "scale_small": iq *= np.float32(1e-17)
    if kind == "scale_big": iq *= np.float32(3e14)
    if kind == "silence":
        a = int(rng.integers(0, n)); b = min(n, a + int(rng.integers(1, n)))
        iq[2 * a: 2 * b] = 0
    if kind == "naninf":
        for _ in range(3):
            k = int(rng.integers(0, 2 * n)); iq[k] = rng.choice([np.nan, np.inf, -np.inf, 3e38])
    if kind == "dc": iq[0::2] += np.float32(0.02)
    want = port.run_iq(iq, rate, thr, pmf, co.MA_CANONICAL)
    # chunking
    mode = rng.choice(["one", "few", "many"])
    if mode == "one": chunks = [n]
    elif mode == "few": chunks = list(rng.integers(1, max(2, n // 2), 8))
    else: chunks = list(rng.integers(1, 4000, 400))
    q = am.msg_queue(); rx = am.rx_path(rate, thr, q, use_pmf=pmf)
    if rng.random() < 0.3: rx._ctx.call("amb_set_option", b"resolver", 1)
    frames = []; pos = 0
    for c in chunks + [n]:
        c = int(min(c, n - pos)); last = pos + c >= n
        rx.process(iq[2 * pos: 2 * (pos + c)], flush=last); frames += rx.frames
        pos += c
        if last: break
    ok = (q.strings() == want.msgs) and ([f.sample_index for f in frames] == [int(x) for x in want.index])
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(rate=rate, pmf=pmf, thr=thr, n=n, nb=nb, sigma=sigma, kind=str(kind), mode=str(mode)),
              "oracle", len(want.index), len(want.msgs), "cuda", len(frames), len(q.strings()))
        a, b = set(int(x) for x in want.index), set(f.sample_index for f in frames)
        print("   det missing", sorted(a - b)[:5], "extra", sorted(b - a)[:5])
    rx.close()
    # the same stream time-sharded over 2..5 contexts with random cuts (amb_seek / amb_resolve)
    if n > 60_000 and rng.random() < 0.6:
        from gr_air_modes_b200 import shard
        g = am.query_geometry(rate, thr, pmf)
        k = int(rng.integers(1, 5))
        lo, hi = g.shard_back + 600, n - g.shard_fwd - 8
        cuts = sorted(set(int(x) for x in rng.integers(lo, hi, k)))
        cuts = [c for i, c in enumerate(cuts) if i == 0 or c - cuts[i - 1] > 16]
        plan = shard.time_shard_plan(n, len(cuts) + 1, g, boundaries=cuts)
        q2 = am.msg_queue(); rxs = [am.rx_path(rate, thr, q2, use_pmf=pmf) for _ in plan]
        for r, sp in zip(rxs, plan):
            r.defer_resolve(True); r.seek(sp.first_sample, sp.first_decision)
            r.process(iq[2 * sp.first_sample: 2 * sp.end], flush=sp.flush, collect=False)
        fr2, state, queued = [], (0, 0), 0
        for r, sp in zip(rxs, plan):
            r.resolve(state)
            if not sp.flush: state = r.walk_state()
            r._slicer._first = queued == 0
            queued += r.drain(); fr2 += r.frames; r.close()
        if q2.strings() != want.msgs or [f.sample_index for f in fr2] != [int(x) for x in want.index]:
            bad += 1
            print("MISMATCH (time-shard) case", case, dict(rate=rate, pmf=pmf, thr=thr, n=n, kind=str(kind)), "cuts", cuts)
print("stress: %d cases, %d mismatches, %.1fs" % (ncase, bad, time.time() - t0))
