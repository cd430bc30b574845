"""Short single-GPU workload for ncu: device-resident noise + bursts, a few amb_process calls."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gr_air_modes_b200 as am
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 4e6
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << logn
g = torch.Generator(device="cuda"); g.manual_seed(1)
iq = torch.randn(2 * n, device="cuda", generator=g) * 0.01
q = am.msg_queue(); rx = am.rx_path(rate, 7.0, q, use_pmf=True)
for it in range(iters):
    rx.reset()
    rx.process(iq, flush=True)
print("done", len(rx.frames))
