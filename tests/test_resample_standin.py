"""tools/resample_standin.py (the documented stand-in for radio.py:49-53's 2 -> 4 Msps resampler).
CPU only: chunk invariance, and that a 2 / 2.4 Msps scene resampled to 4 Msps decodes through the oracle at 4 Msps."""
import os
import sys

import numpy as np
import pytest

from gr_air_modes_b200 import synth
from oracle import cpu_oracle as co

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
pytest.importorskip("scipy")


@pytest.mark.parametrize("rate", [2e6, 2.4e6, 3.2e6])
def test_chunking_does_not_change_the_output(rate):
    from resample_standin import StreamResampler
    rng = np.random.default_rng(1)
    x = rng.standard_normal(2 * 50_000).astype(np.float32)
    r = StreamResampler(rate)
    one = r.push(x, last=True)
    assert abs(one.size / 2 - 50_000 * r.up / r.down) <= 1
    for plan in ([7, 1, 6000, 13, 20000], [1000] * 60, [49_999]):
        r = StreamResampler(rate)
        parts, pos = [], 0
        for c in plan + [50_000]:
            c = min(c, 50_000 - pos)
            parts.append(r.push(x[2 * pos: 2 * (pos + c)], last=(pos + c >= 50_000)))
            pos += c
            if pos >= 50_000:
                break
        got = np.concatenate(parts)
        assert got.size == one.size and np.array_equal(got, one)


@pytest.mark.parametrize("rate,min_share", [(2e6, 0.55), (2.4e6, 0.9), (3.2e6, 0.9)])
def test_resampled_scene_decodes_at_4msps(port, rate, min_share):
    """Why radio.py resamples: with a fractional number of samples per chip the chain (int(spc) correlator,
    preamble_impl.cc:90-98,150) loses most bursts at the native rate (5 of 40 at 2.4 Msps here) and nearly none after
    resampling to 4 Msps. At 2 Msps (one sample per chip, bursts at fractional offsets) both ways lose some."""
    from resample_standin import StreamResampler
    sc = synth.make_scene(rate, int(0.1 * rate), 40, 9, snr_db=(14.0, 30.0), min_gap=200.0)
    r = StreamResampler(rate)
    y = r.push(sc.iq, last=True)
    assert abs(r.rate_out - 4e6) < 1.0
    got = {m.split()[0] for m in port.run_iq(y, 4e6, 7.0, True, co.MA_CANONICAL).msgs}
    sent = {b.frame.hex() for b in sc.bursts}
    native = {m.split()[0] for m in port.run_iq(sc.iq, rate, 7.0, True, co.MA_CANONICAL).msgs}
    assert len(got & sent) >= min_share * len(sent)
    if rate != 2e6:
        assert len(got & sent) > 2 * len(native & sent)
