"""Batch field decode on the GPU: the host-side mirror of what the reference does per message
in Python after the slicer.

  parse.modes_reply(...) field tables, decode_id            python/parse.py:27-254
  altitude.decode_alt                                       python/altitude.py:28-108
  parse.parseBDS05/06/08/09_0/09_1/09_3/62, parseMB_*       python/parse.py:256-420
  cpr.cpr_decoder(my_location).decode(...)                  python/cpr.py:183-240

`batch_decoder` plays the role of one `cpr_decoder` plus the parser: feed it the frames (or the message strings) of a
stretch of the stream in order, get one record per message (struct amb_fields, include/airmodes_b200.h). All
arithmetic runs in libairmodes_b200.so on the GPU; there is no CPU implementation behind this module.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Fields, Frame, check

FS_NO_HANDLER, FS_METRIC_ALT, FS_CPR_NO_POS, FS_CPR_STRADDLE = 0x01, 0x02, 0x04, 0x08
FS_HAS_POS, FS_HAS_RANGE, FS_METRIC_THREAT, FS_NOT_QUEUED = 0x10, 0x20, 0x40, 0x80
NO_ALTITUDE = -(1 << 31)

FIELDS_DTYPE = np.dtype(Fields)
FRAME_DTYPE = np.dtype(Frame)


def frames_from_messages(msgs):
    """The slicer's message strings "<hex> <crc> <ref> <secs> <frac>" (slicer_impl.cc:186-192) -> Frame array,
    i.e. what make_parser's publish() splits (parse.py:425). Every message counts as queued (passed = 1)."""
    msgs = list(msgs)
    arr = (Frame * max(len(msgs), 1))()
    for k, m in enumerate(msgs):
        if isinstance(m, str):
            data, ecc, ref, secs, frac = m.split()
            ecc, ref, secs, frac = int(ecc, 16), float(ref), int(secs), float(frac)
        else:                                   # (hex, ecc, secs, frac) tuples
            data, ecc, secs, frac = m
            ref = 0.0
        raw = bytes.fromhex(data)
        f = arr[k]
        f.nbits = 8 * len(raw)
        for i, b in enumerate(raw):
            f.data[i] = b
        f.crc, f.ref_level, f.secs, f.frac, f.passed = ecc, ref, secs, frac, 1
        f.df = raw[0] >> 3
    return arr, len(msgs)


class batch_decoder:
    """cpr_decoder(my_location) (cpr.py:184-190) + parser, for batches. my_location = [lat, lon] or None."""

    def __init__(self, my_location=None, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        have, lat, lon = self._loc(my_location)
        rc = self._lib.amb_decoder_create(int(device), have, lat, lon, C.byref(h))
        if rc < 0:
            detail = self._lib.amb_decoder_last_error(None).decode()
            raise RuntimeError("libairmodes_b200: %s%s" % (self._lib.amb_strerror(rc).decode(), (" (%s)" % detail) if detail else ""))
        self._h = h
        self.my_location = my_location

    @staticmethod
    def _loc(loc):
        return (0, 0.0, 0.0) if loc is None else (1, float(loc[0]), float(loc[1]))

    def _check(self, rc):
        if rc < 0:
            detail = self._lib.amb_decoder_last_error(self._h).decode()
            raise RuntimeError("libairmodes_b200: %s%s" % (self._lib.amb_strerror(rc).decode(),
                                                           (" (%s)" % detail) if detail else ""))
        return rc

    def set_location(self, new_location):       # cpr.py:192-193
        self._check(self._lib.amb_decoder_set_location(self._h, *self._loc(new_location)))
        self.my_location = new_location

    def reset(self):
        self._check(self._lib.amb_decoder_reset(self._h))

    def decode(self, frames, n: int | None = None) -> np.ndarray:
        """frames: ctypes Frame array / list of Frame (e.g. rx_path.frames) / numpy FRAME_DTYPE array, in stream
        order. Returns a numpy structured array (FIELDS_DTYPE), one record per frame."""
        if isinstance(frames, np.ndarray):
            a = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
            n = a.size if n is None else n
            ptr, keep = C.c_void_p(a.ctypes.data), a
        else:
            if isinstance(frames, (list, tuple)):
                n = len(frames) if n is None else n
                arr = (Frame * max(n, 1))(*frames[:n])
            else:
                arr = frames
                n = len(arr) if n is None else n
            ptr, keep = C.cast(arr, C.c_void_p), arr
        out = np.zeros(n, dtype=FIELDS_DTYPE)
        if n:
            self._check(self._lib.amb_decode_frames(self._h, ptr, int(n), _lib.MEM_HOST, C.c_void_p(out.ctypes.data)))
        del keep
        return out

    def decode_device(self, frames_dev, out_dev=None):
        """frames_dev: torch CUDA uint8 tensor holding n amb_frame records (80 bytes each) in stream order; returns (or
        fills) a CUDA uint8 tensor of n amb_fields records (144 bytes each). Nothing crosses PCIe."""
        import torch
        n = frames_dev.numel() // 80
        if out_dev is None:
            out_dev = torch.empty(n * 144, dtype=torch.uint8, device=frames_dev.device)
        if n:
            torch.cuda.current_stream(frames_dev.device).synchronize()      # the decoder runs on its own stream
            self._check(self._lib.amb_decode_frames_device(self._h, C.c_void_p(frames_dev.data_ptr()), int(n),
                                                           C.c_void_p(out_dev.data_ptr())))
        return out_dev

    def decode_messages(self, msgs) -> np.ndarray:
        arr, n = frames_from_messages(msgs)
        return self.decode(arr, n)

    def stats(self):
        k, ms = C.c_uint64(), C.c_float()
        self._check(self._lib.amb_decoder_stats(self._h, C.byref(k), C.byref(ms)))
        return int(k.value), float(ms.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.amb_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class cpr_decoder:
    """Drop-in for the reference's `air_modes.cpr_decoder(my_location)` (python/cpr.py:183-240; constructed at
    apps/modes_rx:69, used by parseBDS05/06, parse.py:285,291): the same two methods, the same return list
    `[lat, lon, range, bearing]` and the same exceptions, one message per call - so the reference's unmodified
    output plugins can run on top of the GPU decoder. Reports are aged with `clock()` (default time.time, exactly
    as cpr.py:219-221 does). For throughput use batch_decoder; this class is the compatibility surface."""

    def __init__(self, my_location, device: int = 0, clock=None, _backend=None):
        import time
        self.my_location = my_location
        self._clock = clock or time.time
        self._dec = _backend if _backend is not None else batch_decoder(my_location, device)
        self._arr = (Frame * 1)()

    def set_location(self, new_location):                       # cpr.py:192-193
        self.my_location = new_location
        self._dec.set_location(new_location)

    def decode(self, icao24, encoded_lat, encoded_lon, cpr_format, surface):
        from .errors import CPRBoundaryStraddleError, CPRNoPositionError
        now = float(self._clock())
        f = self._arr[0]
        # a DF17 extended squitter carrying exactly this report: ME type 11 (airborne) / 6 (surface), parse.py:124-127
        me = ((6 if surface else 11) << 51) | ((1 if cpr_format else 0) << 34) | ((int(encoded_lat) & 0x1FFFF) << 17) \
            | (int(encoded_lon) & 0x1FFFF)
        raw = bytes([17 << 3]) + (int(icao24) & 0xFFFFFF).to_bytes(3, "big") + me.to_bytes(7, "big") + b"\0\0\0"
        for i, b in enumerate(raw):
            f.data[i] = b
        f.nbits, f.df, f.passed, f.crc = 112, 17, 1, 0
        f.secs = int(now)
        f.frac = now - int(now)
        r = self._dec.decode(self._arr, 1)[0]
        st = int(r["status"])
        if st & FS_CPR_STRADDLE:
            raise CPRBoundaryStraddleError
        if not (st & FS_HAS_POS):
            raise CPRNoPositionError
        if st & FS_HAS_RANGE:
            return [float(r["lat"]), float(r["lon"]), float(r["range"]), float(r["bearing"])]
        return [float(r["lat"]), float(r["lon"]), None, None]

    def close(self):
        if hasattr(self._dec, "close"):
            self._dec.close()


def record_to_dict(rec) -> dict:
    """One FIELDS_DTYPE record -> plain dict (ident as str, arrays as lists)."""
    d = {}
    for name in FIELDS_DTYPE.names:
        v = rec[name]
        if name == "ident":
            v = np.asarray(v).tobytes().rstrip(b"\0").decode("ascii")
        elif isinstance(v, np.ndarray):
            v = v.tolist()
        else:
            v = v.item()
        d[name] = v
    return d


def frame_bits(frame: Frame, start: int, num: int) -> int:
    """data_field.get_bits(start, num) on a frame's payload (parse.py:71-87), for raw sub-fields."""
    return int(_lib.load().amb_frame_bits(C.byref(frame), int(start), int(num)))
