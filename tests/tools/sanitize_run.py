"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co
port = co.Port()
for rate, pmf, dc, dense in ((4e6, True, False, None), (10e6, True, False, None), (5e6, True, False, 0), (2e6, False, False, None),
                             (4e6, True, True, None), (4e6, True, False, 0), (20e6, True, False, 0), (20e6, True, True, None)):
    sc = synth.make_scene(rate, 200_000, 20, 5)
    want = port.run_iq(sc.iq, rate, 7.0, pmf, co.MA_CANONICAL, use_dcblock=dc).msgs
    q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=pmf, use_dcblock=dc)
    rx.set_option("ingest_chunk", 32768)          # several wrap-arounds of the host ingest ring
    if dense is not None:
        rx.set_option("exact_dense", dense)       # the row-based exact kernel (dense-traffic regime)
    for k in range(0, 200_000, 50_000):
        rx.process(sc.iq[2 * k: 2 * (k + 50_000)], flush=(k + 50_000 >= 200_000), collect=False)
    rx.drain()
    print(rate, pmf, dc, dense, q.strings() == want, len(want))
    rx.close()
# 16-bit IQ through the widening kernel, small calls with the non-blocking poll
sc = synth.make_scene(4e6, 150_000, 15, 6)
i16 = np.clip(np.rint(sc.iq * 32768.0), -32768, 32767).astype(np.int16)
want = port.run_iq(i16.astype(np.float32) * np.float32(1 / 32768.0), 4e6, 7.0, True, co.MA_CANONICAL).msgs
q = am.msg_queue(); rx = am.rx_path(4e6, 7.0, q, use_pmf=True)
rx.set_option("coalesce", 20_000)
for k in range(0, 150_000, 8192):
    rx.process(i16[2 * k: 2 * min(k + 8192, 150_000)], collect=False); rx.poll_ready()
rx.process(i16[:0], flush=True)
print("sc16 small calls", q.strings() == want, len(want)); rx.close()
bb, avg = port.frontend(sc.iq, 4e6, True, co.MA_CANONICAL)
pre = am.preamble(4e6, 7.0); chips, tags = pre.process(bb, avg); print("split", len(tags))
# device-side drain (ordering network incl. its global stages: tile 16) -> decoder on device memory
import torch
from gr_air_modes_b200 import decode
sc = synth.make_scene(4e6, 300_000, 160, 7, garble_frac=0.2)
q = am.msg_queue(); rx = am.rx_path(4e6, 7.0, q, use_pmf=True)
rx.set_option("order_tile", 16); rx.add_time_tag(100_000, 50, 0.5)
dev = torch.from_numpy(sc.iq).cuda()
for k in range(0, 300_000, 100_000):
    rx.process(dev[2 * k: 2 * (k + 100_000)], flush=(k + 100_000 >= 300_000), collect=False)
fr = rx.drain_device(); d = decode.batch_decoder([40.0, -3.0]); out = d.decode_device(fr); torch.cuda.synchronize()
print("device drain + decode", fr.numel() // 80, out.numel() // 144); d.close(); rx.close()
