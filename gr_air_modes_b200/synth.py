"""Seeded synthetic 1090 MHz Mode S IQ scenes.

Tooling for tests and bench.py, not part of the receive path. The waveform model follows what the
reference detector/slicer expect: 0.5 us chips at 2 Mchip/s (preamble_impl.cc:46), preamble pulses
at chips 0, 2, 7, 9 (preamble_impl.cc:88-90, 158-162), PPM data from chip 16 with bit 1 = (on, off)
(slicer_impl.cc:133-151), 56/112 data bits, parity = Mode S CRC with polynomial 0xFFF409
(modes_crc.cc:31).  Output format is interleaved float32 I/Q = gr_complex (rx_path.py:29).
"""
from __future__ import annotations

import dataclasses
import hashlib

import numpy as np

POLY = 0xFFF409


def crc24(data: bytes) -> int:
    """Bitwise Mode S CRC over `data` (same polynomial/init as modes_crc.cc:31-63), tool-local."""
    crc = 0
    for byte in data:
        crc ^= byte << 16
        for _ in range(8):
            crc <<= 1
            if crc & 0x1000000:
                crc ^= 0x1000000 | POLY
    return crc & 0xFFFFFF


def make_frame(df: int, rng: np.random.Generator, garble_bits: int = 0) -> bytes:
    """Random DF11 (56 bit) or DF17-style (112 bit) frame whose last 24 bits are the CRC (II=0)."""
    nbytes = 14 if df in (16, 17, 20, 21) else 7
    body = bytearray(rng.integers(0, 256, nbytes - 3, dtype=np.uint8).tobytes())
    body[0] = ((df & 0x1F) << 3) | (body[0] & 0x07)
    frame = bytearray(body) + crc24(bytes(body)).to_bytes(3, "big")
    for _ in range(garble_bits):
        b = int(rng.integers(0, nbytes * 8))
        frame[b // 8] ^= 0x80 >> (b % 8)
    return bytes(frame)


def frame_chips(frame: bytes) -> np.ndarray:
    """On/off chip pattern (16 preamble chips + 2 per bit)."""
    bits = np.unpackbits(np.frombuffer(frame, dtype=np.uint8))
    chips = np.zeros(16 + 2 * bits.size, dtype=np.float64)
    chips[[0, 2, 7, 9]] = 1.0
    chips[16 + 2 * np.nonzero(bits)[0]] = 1.0
    chips[17 + 2 * np.nonzero(bits == 0)[0]] = 1.0
    return chips


@dataclasses.dataclass
class Burst:
    start: float          # fractional sample index of the leading edge of the first preamble pulse
    frame: bytes
    amplitude: float
    phase: float
    freq: float = 0.0     # carrier offset, cycles/sample


def burst_waveform(b: Burst, spc: float):
    """(first_sample, complex64 samples): box-car integrated rectangular pulses at a fractional offset."""
    chips = frame_chips(b.frame)
    knots = np.arange(chips.size + 1) * spc
    cum = np.concatenate([[0.0], np.cumsum(chips) * spc])
    n0 = int(np.floor(b.start))
    n1 = int(np.ceil(b.start + chips.size * spc)) + 1
    edges = np.arange(n0, n1 + 1, dtype=np.float64) - b.start
    integ = np.interp(edges, knots, cum)
    cover = np.diff(integ)  # fraction of each sample interval covered by "on"
    n = np.arange(n0, n1, dtype=np.float64)
    ph = b.phase + 2 * np.pi * b.freq * (n - b.start)
    w = (b.amplitude * cover) * np.exp(1j * ph)
    return n0, w.astype(np.complex64)


@dataclasses.dataclass
class Scene:
    rate: float
    n: int
    iq: np.ndarray                # float32, shape (2n,), interleaved I/Q
    bursts: list

    def sha256(self) -> str:
        return hashlib.sha256(self.iq.tobytes()).hexdigest()


def make_scene(rate: float, n: int, n_bursts: int, seed: int, *, noise_sigma: float = 0.01,
               snr_db=(6.0, 30.0), df_choices=(11, 17), garble_frac: float = 0.0,
               starts=None, amplitude=None, quantize_bits: int | None = None,
               fruit: int = 0, min_gap: float | None = None) -> Scene:
    """Complex Gaussian noise (sigma per component) + `n_bursts` Mode S bursts.

    SNR = A^2 / (2 sigma^2) drawn uniformly in dB from `snr_db`; start times uniform with fractional
    offsets (or Poisson-like overlap when min_gap is None); random carrier phase per burst.
    `quantize_bits` rounds I/Q to a signed fixed-point grid (exactly representable, for small fixtures).
    `fruit` adds that many isolated 0.45 us Mode A/C-like pulses pairs.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    spc = rate / 2e6
    iq = np.empty((n, 2), dtype=np.float32)
    step = 1 << 22
    for a in range(0, n, step):  # chunked to bound temporaries
        m = min(step, n - a)
        iq[a:a + m] = rng.standard_normal((m, 2), dtype=np.float32) * np.float32(noise_sigma)
    c = iq.view(np.complex64).reshape(n)
    bursts = []
    span = 240 * spc + 8
    if starts is None:
        if min_gap is not None:
            # non-overlapping: jittered grid
            slots = int((n - 2 * span) // (span + min_gap))
            pick = np.sort(rng.choice(slots, size=min(n_bursts, slots), replace=False))
            starts = span / 2 + pick * (span + min_gap) + rng.uniform(0, min_gap, pick.size)
        else:
            starts = np.sort(rng.uniform(span, n - 2 * span, n_bursts))
    for k, s in enumerate(starts):
        df = int(df_choices[int(rng.integers(0, len(df_choices)))])
        garble = int(rng.integers(1, 6)) if rng.random() < garble_frac else 0
        frame = make_frame(df, rng, garble)
        if amplitude is None:
            snr = rng.uniform(*snr_db)
            amp = float(np.sqrt(2.0 * noise_sigma ** 2 * 10 ** (snr / 10))) if noise_sigma > 0 else 0.5
        else:
            amp = float(amplitude)
        b = Burst(float(s), frame, amp, float(rng.uniform(0, 2 * np.pi)))
        n0, w = burst_waveform(b, spc)
        lo, hi = max(n0, 0), min(n0 + w.size, n)
        if hi > lo:
            c[lo:hi] += w[lo - n0:hi - n0]
        bursts.append(b)
    for _ in range(fruit):
        s = rng.uniform(0, n - 64 * spc)
        amp = float(np.sqrt(2.0 * max(noise_sigma, 1e-3) ** 2 * 10 ** (rng.uniform(10, 30) / 10)))
        ph = rng.uniform(0, 2 * np.pi)
        for off in (0.0, rng.integers(1, 14) * 1.45 * 2 * spc):  # two pulses 1.45 us multiples apart
            a0 = int(s + off)
            a1 = min(a0 + max(int(round(0.9 * spc)), 1), n)
            c[a0:a1] += np.complex64(amp * np.exp(1j * ph))
    if quantize_bits is not None:
        q = float(1 << (quantize_bits - 1))
        np.clip(np.rint(iq * q), -q, q - 1, out=iq)
        iq /= np.float32(q)
    return Scene(rate, n, iq.reshape(2 * n), bursts)
