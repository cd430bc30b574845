"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import gr_air_modes_b200 as am
from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co
port = co.Port()
for rate, pmf, dc in ((4e6, True, False), (10e6, True, False), (5e6, True, False), (2e6, False, False), (4e6, True, True)):
    sc = synth.make_scene(rate, 200_000, 20, 5)
    want = port.run_iq(sc.iq, rate, 7.0, pmf, co.MA_CANONICAL, use_dcblock=dc).msgs
    q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=pmf, use_dcblock=dc)
    for k in range(0, 200_000, 50_000):
        rx.process(sc.iq[2 * k: 2 * (k + 50_000)], flush=(k + 50_000 >= 200_000), collect=False)
    rx.drain()
    print(rate, pmf, dc, q.strings() == want, len(want))
    rx.close()
bb, avg = port.frontend(sc.iq, 4e6, True, co.MA_CANONICAL)
pre = am.preamble(4e6, 7.0); chips, tags = pre.process(bb, avg); print("split", len(tags))
