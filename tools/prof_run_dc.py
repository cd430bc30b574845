"""Short single-GPU workload for ncu: the optional DC blocker stage (rx_path use_dcblock=True)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gr_air_modes_b200 as am
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 4e6
n = 1 << logn
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randn(2 * n, device="cuda", generator=g) * 0.01 + 0.02
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True, use_dcblock=True)
for it in range(2):
    rx.reset()
    rx.process(iq, flush=True)
print("done", len(rx.frames))
