"""Stand-in for the resampler python/radio.py:49-53 puts in front of rx_path when the source rate is below 4 Msps
(`pfb.arb_resampler_ccf(4.e6 / rate)`), for tools/modes_rx_b200.py --resample.

GNU Radio's polyphase arbitrary resampler (32 phases, Blackman-Harris low_pass_2 taps, derivative-filter interpolation)
is GNU Radio code and is not in the reference tree, so it cannot be reproduced sample for sample. This stand-in does the
same job - band-limited interpolation to 4 Msps - as a rational polyphase resampler (scipy.signal.resample_poly, Kaiser
window, ratio approximated with a denominator <= 64: 2 -> 4 Msps is exactly x2, 2.4 -> 4 is 5/3, 3.2 -> 4 is 5/4).
It is host-side pre-processing outside the hot path, exactly where the reference runs its resampler; the receive
chain itself still runs on the GPU at 4 Msps. Streaming: any chunking of the input gives bit-identical output
(chunks are cut on multiples of the decimation factor with the filter's support as context on both sides).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np
from scipy.signal import resample_poly


class StreamResampler:
    def __init__(self, rate_in: float, rate_out: float = 4e6, max_den: int = 64):
        fr = Fraction(rate_out / rate_in).limit_denominator(max_den)
        self.up, self.down = fr.numerator, fr.denominator
        self.rate_out = rate_in * self.up / self.down         # exact output rate of the rational approximation
        self.pad = 64 * self.down                             # > resample_poly's half support (10 * max(up, down) / up inputs)
        self._buf = np.zeros(0, np.complex64)                 # input from absolute index _buf0 on
        self._buf0 = 0
        self._emit = 0                                        # next input index whose outputs are still to be produced

    def push(self, iq_f32: np.ndarray, last: bool = False) -> np.ndarray:
        """Interleaved float32 I/Q in (any length), interleaved float32 I/Q out (whatever can be decided)."""
        x = np.ascontiguousarray(iq_f32, dtype=np.float32).view(np.complex64)
        self._buf = np.concatenate([self._buf, x]) if self._buf.size else x.copy()
        end = self._buf0 + self._buf.size                     # absolute end of the input seen so far
        stop = end if last else max(self._emit, (end - self.pad) // self.down * self.down)
        if stop <= self._emit and not (last and end > self._emit):
            return np.zeros(0, np.float32)
        lp = min(self.pad, self._emit)                        # left context (a multiple of down: both are)
        seg_hi = end if last else stop + self.pad
        seg = self._buf[self._emit - lp - self._buf0: seg_hi - self._buf0]
        y = resample_poly(seg, self.up, self.down).astype(np.complex64)
        o0 = lp * self.up // self.down
        o1 = y.size if last else (lp + stop - self._emit) * self.up // self.down
        out = y[o0:o1]
        self._emit = stop
        keep_from = max(self._emit - self.pad, self._buf0)
        self._buf = self._buf[keep_from - self._buf0:]
        self._buf0 = keep_from
        return np.ascontiguousarray(out).view(np.float32)
