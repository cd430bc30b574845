"""Generates the committed golden fixtures. Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

For each scene: a seeded synthetic IQ snippet, quantised to a 12-bit grid and stored as int16 (so the
float32 values are exactly reproducible anywhere), plus what the UNMODIFIED reference
(lib/preamble_impl.cc + lib/slicer_impl.cc + lib/modes_crc.cc, compiled by oracle/Makefile into
oracle/_ref) produces for it behind the canonical front end: detection indices, message strings, and a
SHA-256 of the 240-chip packets. Nothing here is hand-edited.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gr_air_modes_b200 import synth  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402

SCENES = [
    # name, rate, n, bursts, seed, use_pmf, threshold_db, extra
    ("s2msps", 2e6, 1 << 16, 10, 101, True, 7.0, {}),
    ("s4msps", 4e6, 1 << 16, 10, 102, True, 7.0, {}),
    ("s4msps_nopmf_t5", 4e6, 1 << 16, 10, 103, False, 5.0, {}),
    ("s10msps", 10e6, 1 << 17, 10, 104, True, 7.0, {}),
    ("s20msps", 20e6, 1 << 17, 6, 105, True, 7.0, {}),
    ("s5msps_frac", 5e6, 1 << 16, 8, 106, True, 7.0, {}),
    ("s4msps_dense", 4e6, 1 << 16, 60, 107, True, 6.0, {"garble_frac": 0.3, "fruit": 40}),
]
KNOWN = ["8D4840D6202CC371C32CE0576098", "8D40621D58C382D690C8AC2863A7"]


def main():
    port, ref = co.Port(), co.Ref()
    meta = {"scenes": [], "crc": []}
    for hexs in KNOWN:
        b = bytes.fromhex(hexs)
        meta["crc"].append({"frame": hexs, "crc_first_11": "%06x" % ref.crc24(b[:11]),
                            "syndrome": "%06x" % (ref.crc24(b[:11]) ^ int.from_bytes(b[11:], "big"))})
    arrays = {}
    for name, rate, n, nb, seed, pmf, thr, extra in SCENES:
        sc = synth.make_scene(rate, n, nb, seed, quantize_bits=12, noise_sigma=0.02, snr_db=(8.0, 30.0), **extra)
        i16 = np.rint(sc.iq * 2048.0).astype(np.int16)
        assert np.array_equal(i16.astype(np.float32) / np.float32(2048.0), sc.iq)
        arrays[name] = i16
        sc.iq = i16.astype(np.float32) / np.float32(2048.0)      # exactly what the tests reconstruct (no -0.0)
        bb, avg = port.frontend(sc.iq, rate, pmf, co.MA_CANONICAL)
        r = ref.run_streams(bb, avg, rate, thr)
        meta["scenes"].append({
            "name": name, "rate": rate, "n": n, "use_pmf": pmf, "threshold_db": thr, "seed": seed,
            "iq_sha256": hashlib.sha256(sc.iq.tobytes()).hexdigest(),
            "det_index": [int(x) for x in r.index],
            "chips_sha256": hashlib.sha256(np.ascontiguousarray(r.chips).tobytes()).hexdigest(),
            "msgs": r.msgs,
            "sent": [b.frame.hex() for b in sc.bursts],
        })
        print(name, "det", len(r.index), "msgs", len(r.msgs))
    np.savez_compressed(os.path.join(HERE, "scenes.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
