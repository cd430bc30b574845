#!/usr/bin/env python
"""modes_rx-style front end for the B200 receive chain (minimal): read a cfile
(interleaved float32 I/Q, what `modes_rx -s file.cfile` feeds to rx_path, radio.py:221-232) and print the
slicer messages "<hex> <crc> <ref> <secs> <frac>" exactly as they would be put on the msg_queue.

    python tools/modes_rx_b200.py -s capture.cfile -r 4e6 [-T 7.0] [--no-pmf] [--chunk 16777216]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gr_air_modes_b200 as air_modes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", "--source", required=True, help="cfile (complex64)")              # radio.py:94-95
    ap.add_argument("-r", "--rate", type=float, default=4e6)                                # radio.py:112
    ap.add_argument("-T", "--threshold", type=float, default=7.0)                           # radio.py:114
    ap.add_argument("--no-pmf", action="store_true", help="disable the pulse matched filter (CLI default is on, radio.py:116)")
    ap.add_argument("--chunk", type=int, default=1 << 24, help="complex samples per amb_process call")
    args = ap.parse_args()
    q = air_modes.msg_queue()
    rx = air_modes.rx_path(args.rate, args.threshold, q, use_pmf=not args.no_pmf)
    n_total = os.path.getsize(args.source) // 8
    mm = np.memmap(args.source, dtype=np.float32, mode="r", shape=(2 * n_total,))
    pos = 0
    while pos < n_total or n_total == 0:
        c = min(args.chunk, n_total - pos)
        rx.process(np.asarray(mm[2 * pos: 2 * (pos + c)]), flush=(pos + c >= n_total))
        pos += c
        while not q.empty_p():
            print(q.delete_head().to_string())
        if n_total == 0:
            break
    st = rx.stats()
    print("# %d samples, %d messages-capable detections in the last call" % (st.samples_in, st.detections), file=sys.stderr)


if __name__ == "__main__":
    main()
