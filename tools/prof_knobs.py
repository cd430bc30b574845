"""Run-time knobs measured side by side on one box (no rebuild needed):
  A  which exact-stage regime (option exact_dense) is faster per BASELINE config - warp per candidate or rows;
  B  pageable host input: copy threads x ingest chunk size.
    python tools/prof_knobs.py [A] [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import gr_air_modes_b200 as am

what = [a for a in sys.argv[1:] if a in ("A", "B")] or ["A", "B"]
dev = torch.device("cuda")
n = 1 << 28


def step_ms(cfg, iq, opts, K=20):
    q = am.msg_queue(); rx = am.rx_path(cfg["rate"], 7.0, q, use_pmf=True)
    for k, v in opts.items():
        rx.set_option(k, v)
    rx._ctx.call("amb_enable_timing", 1)
    rx.use_stream(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    rx.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(K):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    rx.join(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    sc = float(np.mean(rx._ctx.scan_times_ms(K)))
    nm = rx.drain()
    rx.close()
    return ms, sc, nm


if "A" in what:
    for key in ("c1", "c2", "c3"):
        cfg = bench.CONFIGS[key]
        iq, _ = bench.make_device_scene(cfg, n, 0, dev)
        torch.cuda.synchronize()
        for name, opts in (("warp per candidate", {}), ("rows", {"exact_dense": 0})):
            r = [step_ms(cfg, iq, opts) for _ in range(2)]
            print("A %s exact stage = %-18s step %.4f / %.4f ms  scan %.4f ms  msgs %d" % (key, name, r[0][0], r[1][0], r[1][1], r[1][2]), flush=True)
        del iq

if "B" in what:
    cfg = bench.CONFIGS["c1"]
    bench.bind_near_gpu(0)
    iq, _ = bench.make_device_scene(cfg, n, 0, dev)
    page = np.empty(2 * n, dtype=np.float32)
    page[:] = iq.cpu().numpy()
    pinned = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
    pinned.copy_(iq); del iq
    torch.cuda.synchronize()

    def timed(rx, buf, reps=3):
        rx.reset(); rx.process(buf, flush=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            rx.reset(); rx.process(buf, flush=True)
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / reps

    q = am.msg_queue(); rx = am.rx_path(cfg["rate"], 7.0, q, use_pmf=True)
    base = timed(rx, pinned)
    print("B pinned float32: %.2f ms = %.2f GS/s" % (base, n / base / 1e6), flush=True)
    rx.close()
    for chunk in (1 << 21, 1 << 22, 1 << 23):
        for thr in (0, 8, 12, 16, 24, 32, 48):
            q = am.msg_queue(); rx = am.rx_path(cfg["rate"], 7.0, q, use_pmf=True)
            rx.set_option("ingest_chunk", chunk)
            rx.set_option("copy_threads", thr)
            ms = timed(rx, page)
            print("B pageable chunk 2^%d threads %2d: %.2f ms = %.2f GS/s = %.0f %% of pinned" % (
                chunk.bit_length() - 1, thr, ms, n / ms / 1e6, 100 * base / ms), flush=True)
            rx.close()
