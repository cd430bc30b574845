import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        meta = json.load(f)
    z = np.load(os.path.join(HERE, "golden", "scenes.npz"))
    scenes = []
    for s in meta["scenes"]:
        iq = z[s["name"]].astype(np.float32) / np.float32(2048.0)
        scenes.append((s, iq))
    return meta, scenes


def parse_msg(m):
    """'<hex> <crc> <ref> <secs> <frac>' (slicer_impl.cc:186-192, parse.py:425)."""
    data, crc, ref, secs, frac = m.split()
    return data, int(crc, 16), float(ref), int(secs), float(frac)


def first_four(bb, avg, rate, threshold_db, port):
    """Indices (reported coordinates) passing preamble_impl.cc:173-179, vectorised in float32."""
    p = port.params(rate, threshold_db)
    H = p.history - 1
    n = bb.size
    pad = 16 * int(np.ceil(p.spc)) + 8
    a = np.concatenate([np.zeros(H, np.float32), bb, np.zeros(pad, np.float32)])
    v = np.concatenate([np.zeros(H, np.float32), avg, np.zeros(pad, np.float32)])
    m = n + H
    thr = (v[:m] * np.float32(p.threshold)).astype(np.float32)
    ok = a[:m] > thr
    ok &= ~(a[1:m + 1] > a[:m])
    for k in (1, 2, 3):
        ok &= ~(a[p.po[k]:m + p.po[k]] < thr)
    return np.nonzero(ok)[0]
