#!/usr/bin/env python
"""Timing / profiling driver for the batch field decoder: synthetic DF17 position traffic,
vectorised frame construction (no torch import), decode at a few batch sizes.

    python tests/tools/prof_decode.py [--check] [log2n ...]   # device ms per batch (H2D + 3 kernels + D2H); --check: parity first
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/decode_launches.csv python tests/tools/prof_decode.py 16 20
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gr_air_modes_b200 import decode  # noqa: E402


def make_frames(n, n_aircraft, seed):
    rng = np.random.default_rng(seed)
    f = np.zeros(n, dtype=decode.FRAME_DTYPE)
    icao = rng.choice(1 << 24, n_aircraft, replace=False)[rng.integers(0, n_aircraft, n)]
    data = rng.integers(0, 256, (n, 14), dtype=np.uint8)
    data[:, 0] = 0x8D                                   # DF17, CA 5
    data[:, 1], data[:, 2], data[:, 3] = (icao >> 16) & 0xFF, (icao >> 8) & 0xFF, icao & 0xFF
    data[:, 4] = (11 << 3) | (data[:, 4] & 0x07)        # ME type 11: airborne position, the rest random
    f["data"] = data
    f["nbits"], f["df"], f["passed"] = 112, 17, 1
    t = np.cumsum(rng.exponential(1e-4, n))             # 10 k messages/s
    f["secs"] = t.astype(np.uint64)
    f["frac"] = t - np.floor(t)
    return f


def check():
    """Parity of whatever libairmodes_b200.so is in place against the CPU oracle on a seeded case (tests/decode_cases)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import decode_cases
    from oracle import decode_oracle as do
    bad = 0
    for seed, loc in ((501, (35.7, 139.7)), (502, None)):
        loc_, msgs = decode_cases.make_case(seed, location=loc, seconds=30.0, surface_share=0.4, n_random=600, n_aircraft=12)
        want = do.decode_batch(msgs, loc_)
        d = decode.batch_decoder(loc_)
        got = d.decode_messages(msgs)
        d.close()
        for g, w in zip(got, want):
            same = g["status"] == w["status"] and g["altitude"] == w["altitude"]
            if w["status"] & do.FS_HAS_POS:
                same = same and abs(g["lat"] - w["lat"]) < 1e-10 and abs(g["lon"] - w["lon"]) < 1e-10
            bad += not same
    print("parity vs oracle: %s" % ("ok" if bad == 0 else "%d MISMATCHES" % bad))
    return bad == 0


def main():
    args = [a for a in sys.argv[1:] if a != "--check"]
    if "--check" in sys.argv[1:] and not check():
        sys.exit(1)
    sizes = [int(a) for a in args] or [16, 20]
    d = decode.batch_decoder([40.0, -3.0])
    for lg in sizes:
        n = 1 << lg
        f = make_frames(n, max(50, n // 200), lg)
        d.reset()
        d.decode(f[:1024])                              # warm-up (buffers, module load)
        d.reset()
        t0 = time.perf_counter()
        out = d.decode(f)
        wall = time.perf_counter() - t0
        ms = d.stats()[1]
        npos = int(((out["status"] & decode.FS_HAS_POS) != 0).sum())
        print("n=2^%d frames: device %.3f ms (%.1f M frames/s), wall %.1f ms incl. numpy; %d positions, %d straddles/no-pair"
              % (lg, ms, n / ms / 1e3, 1e3 * wall, npos, n - npos))
        try:                                            # the same batch with frames and records resident on the device
            import torch
            fd = torch.from_numpy(f.view(np.uint8).reshape(-1).copy()).cuda()
            od = torch.empty(n * 144, dtype=torch.uint8, device="cuda")
            d.reset(); d.decode_device(fd, od)
            d.reset(); d.decode_device(fd, od)
            ms2 = d.stats()[1]
            same = od.cpu().numpy().tobytes() == out.tobytes()        # bytes: the records hold NaN for "no position"
            print("n=2^%d frames, device-resident frames and records: %.3f ms (%.0f M frames/s), records identical: %s"
                  % (lg, ms2, n / ms2 / 1e3, same))
        except ImportError:
            pass
    d.close()


if __name__ == "__main__":
    main()
