"""Row f4 behind the hot path, two ways, on the BASELINE scenes (device-resident IQ, one GPU):
  host-driven : rx_path.drain() (frames D2H, sorted + stamped on the host) -> batch_decoder.decode() (H2D, kernels, D2H)
  on the device: rx_path.drain_device() (ordered + stamped by kernels)     -> batch_decoder.decode_device()
    python tools/prof_chain.py [c1 c4] [log2n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import gr_air_modes_b200 as am
from gr_air_modes_b200 import decode

keys = [a for a in sys.argv[1:] if a in bench.CONFIGS] or ["c1", "c4"]
logn = [int(a) for a in sys.argv[1:] if a.isdigit()]
n = 1 << (logn[0] if logn else 28)
dev = torch.device("cuda")


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); t.append(1e3 * (time.perf_counter() - t0))
    return min(t), r


for key in keys:
    cfg = bench.CONFIGS[key]
    iq, _ = bench.make_device_scene(cfg, n, 0, dev)
    torch.cuda.synchronize()
    rx = am.rx_path(cfg["rate"], 7.0, am.msg_queue(), use_pmf=True)
    dh, dd = decode.batch_decoder([40.0, -3.0]), decode.batch_decoder([40.0, -3.0])

    def scan():
        rx.reset(); rx.process(iq, flush=True, collect=False); rx._ctx.call("amb_synchronize")

    def host_drain():
        scan(); t0 = time.perf_counter(); buf, got = rx._ctx.poll_array(); return 1e3 * (time.perf_counter() - t0), buf, got

    def dev_drain():
        scan(); t0 = time.perf_counter(); fr = rx.drain_device(); return 1e3 * (time.perf_counter() - t0), fr

    host_drain(); dev_drain()
    th, buf, got = min((host_drain() for _ in range(3)), key=lambda x: x[0])
    td, fr = min((dev_drain() for _ in range(3)), key=lambda x: x[0])
    same = fr.cpu().numpy().tobytes() == bytes(buf)[:got * 80]
    arr = np.frombuffer(bytes(buf)[:got * 80], dtype=decode.FRAME_DTYPE)
    dh.decode(arr[:1024]); dd.decode_device(fr[:80 * min(got, 1024)])

    def host_dec():
        dh.reset(); return dh.decode(arr)

    def dev_dec():
        dd.reset(); return dd.decode_device(fr)

    tdh, want = wall(host_dec)
    tdd, out = wall(dev_dec)
    same_rec = out.cpu().numpy().tobytes() == want.tobytes()
    print("%s 2^%d samples, %d frames: drain host %.3f ms / device %.3f ms (identical %s); decode host arrays %.3f ms / device %.3f ms "
          "(identical %s); chain host %.3f ms, device %.3f ms" % (key, n.bit_length() - 1, got, th, td, same, tdh, tdd, same_rec,
                                                                  th + tdh, td + tdd), flush=True)
    rx.close(); dh.close(); dd.close(); del iq
