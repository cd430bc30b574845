"""Experiment: how many CTAs should the scan kernel use? One full wave (4 per SM) leaves the sparse kernels of the
previous call nowhere to run but the scan's leftovers; a few CTA slots less gives them whole slots.
    python tools/prof_holes.py c1|c2|c3|c4 [log2n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import gr_air_modes_b200 as am
key = sys.argv[1] if len(sys.argv) > 1 else "c1"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 28
cfg = bench.CONFIGS[key]; n = 1 << logn
dev = torch.device("cuda")
iq, _ = bench.make_device_scene(cfg, n, 0, dev)
torch.cuda.synchronize()
sm = torch.cuda.get_device_properties(0).multi_processor_count
for ctas in (4 * sm,):
    q = am.msg_queue(); rx = am.rx_path(cfg["rate"], 7.0, q, use_pmf=True)
    rx.set_option("scan_ctas", ctas)
    rx._ctx.call("amb_enable_timing", 1)
    rx.use_stream(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    rx.join(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record()
    for it in range(K):
        rx.reset(); rx.process(iq, flush=True, collect=False)
    rx.join(); e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / K
    sc = float(np.mean(rx._ctx.scan_times_ms(K)))
    import ctypes as C
    buf = (C.c_float * (3 * 20))()
    k = rx._ctx.call("amb_get_timeline", buf, 20)
    tl = np.array(buf[:3 * k]).reshape(k, 3)[3:]
    nm = rx.drain()
    print("   timeline (mean over calls): scan-stream idle before scan %.4f ms, scan %.4f ms, scan end -> sparse done %.4f ms" % tuple(tl.mean(0)))
    print("%s scan_ctas %4d: step %.4f ms (%.1f GS/s)  scan under overlap %.4f ms  msgs %d" % (key, ctas, b2b, n / b2b / 1e6, sc, nm))
    rx.close()
