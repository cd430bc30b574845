"""Short single-GPU workload for ncu: 16-bit IQ from pinned host memory through the ingest ring (amb_widen_sc16_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gr_air_modes_b200 as am
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << logn
g = torch.Generator(); g.manual_seed(1)
i16 = torch.empty(2 * n, dtype=torch.int16, pin_memory=True)
i16.copy_(torch.randint(-300, 300, (2 * n,), generator=g, dtype=torch.int16))
q = am.msg_queue(); rx = am.rx_path(4e6, 7.0, q, use_pmf=True)
for it in range(2):
    rx.reset()
    rx.process(i16, flush=True)
print("done", len(rx.frames))
