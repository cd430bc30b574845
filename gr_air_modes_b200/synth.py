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


# ---------------------------------------------------------------------------------------------------------
# The same scene model, vectorised (numpy for the frames, torch for the waveforms): bench.py needs scenes of
# 2^28 samples with up to ~10^6 bursts (BASELINE configs[4]), which the per-burst Python loop above cannot
# deliver in bench time. Not bit-identical to make_scene (other random streams); parity is always checked
# against the oracle on the very buffer this returns.
# ---------------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, np.uint32)
        for n in range(256):
            c = n << 16
            for _ in range(8):
                c = ((c << 1) ^ POLY) & 0xFFFFFF if c & 0x800000 else (c << 1) & 0xFFFFFF
            t[n] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def make_frames(dfs: np.ndarray, rng: np.random.Generator, garble_frac: float = 0.0):
    """Vectorised make_frame: (frames uint8[nb,14], nbytes int[nb]); short frames are zero beyond byte 7."""
    nb = dfs.size
    long_ = np.isin(dfs, (16, 17, 20, 21))
    nbytes = np.where(long_, 14, 7)
    fr = rng.integers(0, 256, (nb, 14), dtype=np.uint8)
    fr[:, 0] = ((dfs.astype(np.uint8) & 0x1F) << 3) | (fr[:, 0] & 0x07)
    T = _crc_table()
    crc_s = np.zeros(nb, np.uint32)
    crc_l = np.zeros(nb, np.uint32)
    for i in range(11):
        crc_l = T[((crc_l >> 16) ^ fr[:, i]) & 0xFF] ^ ((crc_l << 8) & 0xFFFFFF)
        if i == 3:
            crc_s = crc_l.copy()
    crc = np.where(long_, crc_l, crc_s)
    col = np.where(long_, 11, 4)
    rows = np.arange(nb)
    for k in range(3):
        fr[rows, col + k] = (crc >> (16 - 8 * k)) & 0xFF
    fr[~long_, 7:] = 0
    if garble_frac > 0:
        g = np.nonzero(rng.random(nb) < garble_frac)[0]
        nflip = rng.integers(1, 6, g.size)
        for k in range(5):
            sel = g[nflip > k]
            bit = (rng.random(sel.size) * (nbytes[sel] * 8)).astype(np.int64)
            fr[sel, bit // 8] ^= (0x80 >> (bit % 8)).astype(np.uint8)
    return fr, nbytes


def make_scene_device(rate: float, n: int, n_bursts: int, seed: int, device, *, noise_sigma: float = 0.01,
                      snr_db=(6.0, 30.0), df_choices=(11, 17), garble_frac: float = 0.0, fruit: int = 0, out=None):
    """(iq float32 CUDA tensor[2n], list of frame hex strings). Same waveform model as make_scene: box-car integrated
    rectangular chips at fractional start offsets, random carrier phase, uniform starts (overlaps allowed)."""
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    spc = rate / 2e6
    g = torch.Generator(device=device)
    g.manual_seed(1000 + seed)
    iq = out if out is not None else torch.empty(2 * n, device=device, dtype=torch.float32)
    step = 1 << 26
    for a in range(0, 2 * n, step):
        m = min(step, 2 * n - a)
        torch.randn(m, device=device, generator=g, out=iq[a:a + m])
        iq[a:a + m] *= noise_sigma
    iq2 = iq.view(n, 2)
    span = 240 * spc + 8
    starts = np.sort(rng.uniform(span, n - 2 * span, n_bursts))
    dfs = np.asarray(df_choices)[rng.integers(0, len(df_choices), n_bursts)]
    frames, nbytes = make_frames(dfs, rng, garble_frac)
    snr = rng.uniform(snr_db[0], snr_db[1], n_bursts)
    amp = np.sqrt(2.0 * noise_sigma ** 2 * 10 ** (snr / 10)) if noise_sigma > 0 else np.full(n_bursts, 0.5)
    phase = rng.uniform(0, 2 * np.pi, n_bursts)
    bits = np.unpackbits(frames, axis=1)                                   # [nb, 112]
    valid = np.arange(112)[None, :] < (nbytes * 8)[:, None]
    chips = np.zeros((n_bursts, 240), np.float32)
    chips[:, [0, 2, 7, 9]] = 1.0
    chips[:, 16:240:2] = bits * valid
    chips[:, 17:240:2] = (1 - bits) * valid
    M = int(np.ceil(240 * spc)) + 2
    mgrid = torch.arange(M + 1, device=device, dtype=torch.float64)
    B = max(1, (1 << 24) // (M + 1))
    for a in range(0, n_bursts, B):
        b = min(a + B, n_bursts)
        s = torch.from_numpy(starts[a:b]).to(device)
        n0 = torch.floor(s)
        ch = torch.from_numpy(chips[a:b]).to(device).to(torch.float64)
        cum = torch.cumsum(ch, 1) - ch                                       # exclusive
        t = ((n0 - s)[:, None] + mgrid[None, :]) / spc                       # edges in chips
        t = t.clamp_(0.0, 240.0)
        k = t.floor().clamp_(max=239.0)
        ki = k.to(torch.int64)
        integ = (torch.gather(cum, 1, ki) + torch.gather(ch, 1, ki) * (t - k)) * spc
        cover = integ[:, 1:] - integ[:, :-1]                                 # [b, M]
        w = torch.from_numpy(amp[a:b]).to(device)[:, None] * cover
        ph = torch.from_numpy(phase[a:b]).to(device)[:, None]
        vals = torch.stack((w * torch.cos(ph), w * torch.sin(ph)), 2).to(torch.float32).reshape(-1, 2)
        idx = (n0.to(torch.int64)[:, None] + torch.arange(M, device=device)[None, :]).reshape(-1)
        ok = (idx >= 0) & (idx < n)
        iq2.index_add_(0, idx[ok], vals[ok])
    if fruit:
        s = rng.uniform(0, n - 64 * spc, fruit)
        fa = np.sqrt(2.0 * max(noise_sigma, 1e-3) ** 2 * 10 ** (rng.uniform(10, 30, fruit) / 10))
        ph = rng.uniform(0, 2 * np.pi, fruit)
        off2 = rng.integers(1, 14, fruit) * 1.45 * 2 * spc
        wlen = max(int(round(0.9 * spc)), 1)
        a0 = np.concatenate([s.astype(np.int64), (s + off2).astype(np.int64)])
        idx = torch.from_numpy((a0[:, None] + np.arange(wlen)[None, :]).reshape(-1)).to(device)
        v = np.stack([fa * np.cos(ph), fa * np.sin(ph)], 1).astype(np.float32)
        vals = torch.from_numpy(np.repeat(np.concatenate([v, v]), wlen, axis=0)).to(device)
        ok = (idx >= 0) & (idx < n)
        iq2.index_add_(0, idx[ok], vals[ok])
    hexs = [frames[k, :nbytes[k]].tobytes().hex() for k in range(n_bursts)] if n_bursts <= 4096 else None
    return iq, hexs, (frames, nbytes)
