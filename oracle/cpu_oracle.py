"""TEST INFRASTRUCTURE ONLY: ctypes access to the CPU oracle.

  Port  - oracle/liboracle_port.so, the plain-C restatement (modes_oracle.c); travels everywhere.
  Ref   - oracle/_ref/libairmodes_ref.so, the UNMODIFIED reference lib/*.cc compiled against
          oracle/shim (built only where /root/reference exists; the prebuilt .so travels to the
          GPU box with the snapshot).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (gr_air_modes_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libairmodes_ref.so")

MA_CANONICAL, MA_GR_FLOAT, MA_SLIDING64 = 0, 1, 2


def build(quiet: bool = True) -> None:
    """Compile the port and, when /root/reference is present, the reference objects."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class Frame(C.Structure):
    _fields_ = [("index", C.c_uint64), ("secs", C.c_uint64), ("frac", C.c_double),
                ("ref_level", C.c_float), ("crc", C.c_uint32), ("nbits", C.c_uint8),
                ("df", C.c_uint8), ("numlowconf", C.c_uint8), ("passed", C.c_uint8),
                ("lowconfbits", C.c_uint8 * 24), ("data", C.c_uint8 * 14)]


class Params(C.Structure):
    _fields_ = [("spc", C.c_float), ("sps", C.c_float), ("check_width", C.c_int),
                ("rate_int", C.c_int), ("history", C.c_int), ("threshold_db", C.c_float),
                ("threshold", C.c_float), ("po", C.c_int * 4)]


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class Result:
    """Detections (+chips), per-detection frames (port only) and queue messages of one run."""

    def __init__(self, index, secs, frac, chips, msgs, frames=None, calls=0):
        self.index, self.secs, self.frac, self.chips = index, secs, frac, chips
        self.msgs, self.frames, self.calls = msgs, frames, calls


class Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        vp, u64, f32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_float)
        L.amo_make_params.argtypes = [C.c_float, C.c_float, C.POINTER(Params)]
        L.amo_set_start_time.argtypes = [C.c_uint64, C.c_double]
        L.amo_crc24.argtypes = [C.c_char_p, C.c_int]; L.amo_crc24.restype = C.c_uint32
        L.amo_mag2.argtypes = [f32p, u64, f32p]
        L.amo_moving_average.argtypes = [f32p, u64, C.c_int, C.c_float, C.c_int, C.c_int, f32p]
        L.amo_frontend.argtypes = [f32p, u64, C.c_float, C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.amo_dc_blocker.argtypes = [f32p, u64, C.c_int, C.c_int, f32p]
        for name, args in (("amo_scan", [f32p, f32p, u64, C.c_float, C.c_float]),
                           ("amo_run_streams", [f32p, f32p, u64, C.c_float, C.c_float]),
                           ("amo_run_iq", [f32p, u64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]),
                           ("amo_run_iq_dc", [f32p, u64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
                           ("amo_run_slicer", [f32p, u64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)])):
            fn = getattr(L, name); fn.argtypes = args; fn.restype = vp
        L.amo_num_det.argtypes = [vp]; L.amo_num_det.restype = u64
        L.amo_num_calls.argtypes = [vp]; L.amo_num_calls.restype = u64
        L.amo_get_det.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double), f32p]
        L.amo_get_frames.argtypes = [vp, C.POINTER(Frame)]
        L.amo_num_msgs.argtypes = [vp]; L.amo_num_msgs.restype = u64
        L.amo_msg.argtypes = [vp, u64]; L.amo_msg.restype = C.c_char_p
        L.amo_free.argtypes = [vp]
        L.amo_slice_packet.argtypes = [f32p, C.POINTER(Frame)]
        L.amo_format_message.argtypes = [C.POINTER(Frame), C.c_int, C.c_char_p, C.c_size_t]
        L.amo_format_message.restype = C.c_int
        self.L = L

    # -- helpers
    def set_start_time(self, secs: int, frac: float):
        self.L.amo_set_start_time(int(secs), float(frac))

    def params(self, rate, threshold_db) -> Params:
        p = Params(); self.L.amo_make_params(rate, threshold_db, C.byref(p)); return p

    def crc24(self, data: bytes) -> int:
        return int(self.L.amo_crc24(bytes(data), len(data)))

    def mag2(self, iq):
        iq, p = _f32(iq); n = iq.size // 2
        out = np.empty(n, np.float32)
        self.L.amo_mag2(p, n, out.ctypes.data_as(C.POINTER(C.c_float))); return out

    def moving_average(self, u, length, scale, mode=MA_CANONICAL, chunk=4096):
        u, p = _f32(u); out = np.empty(u.size, np.float32)
        self.L.amo_moving_average(p, u.size, length, scale, mode, chunk,
                                  out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def frontend(self, iq, rate, use_pmf=True, ma_mode=MA_CANONICAL, chunk=4096):
        iq, p = _f32(iq); n = iq.size // 2
        bb = np.empty(n, np.float32); avg = np.empty(n, np.float32)
        self.L.amo_frontend(p, n, rate, int(use_pmf), ma_mode, chunk,
                            bb.ctypes.data_as(C.POINTER(C.c_float)), avg.ctypes.data_as(C.POINTER(C.c_float)))
        return bb, avg

    def _collect(self, h, with_frames=True) -> Result:
        L = self.L
        n = int(L.amo_num_det(h))
        index = np.empty(n, np.uint64); secs = np.empty(n, np.uint64)
        frac = np.empty(n, np.float64); chips = np.empty((n, 240), np.float32)
        L.amo_get_det(h, index.ctypes.data_as(C.POINTER(C.c_uint64)), secs.ctypes.data_as(C.POINTER(C.c_uint64)),
                      frac.ctypes.data_as(C.POINTER(C.c_double)), chips.ctypes.data_as(C.POINTER(C.c_float)))
        frames = None
        if with_frames:
            frames = (Frame * max(n, 1))()
            L.amo_get_frames(h, frames)
            frames = list(frames)[:n]
        msgs = [L.amo_msg(h, k).decode() for k in range(int(L.amo_num_msgs(h)))]
        calls = int(L.amo_num_calls(h))
        L.amo_free(h)
        return Result(index, secs, frac, chips, msgs, frames, calls)

    def scan(self, bb, avg, rate, threshold_db) -> Result:
        bb, pb = _f32(bb); avg, pa = _f32(avg)
        return self._collect(self.L.amo_scan(pb, pa, bb.size, rate, threshold_db), with_frames=False)

    def run_streams(self, bb, avg, rate, threshold_db) -> Result:
        bb, pb = _f32(bb); avg, pa = _f32(avg)
        return self._collect(self.L.amo_run_streams(pb, pa, bb.size, rate, threshold_db))

    def run_iq(self, iq, rate, threshold_db, use_pmf=True, ma_mode=MA_CANONICAL, chunk=4096, use_dcblock=False) -> Result:
        iq, p = _f32(iq)
        return self._collect(self.L.amo_run_iq_dc(p, iq.size // 2, rate, threshold_db, int(use_pmf), int(use_dcblock),
                                                  ma_mode, chunk))

    def dc_blocker(self, iq, D, mode=MA_CANONICAL):
        iq, p = _f32(iq); out = np.empty(iq.size, np.float32)
        self.L.amo_dc_blocker(p, iq.size // 2, int(D), mode, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def run_slicer(self, chips, secs, frac) -> Result:
        chips, pc = _f32(chips)
        secs = np.ascontiguousarray(secs, np.uint64); frac = np.ascontiguousarray(frac, np.float64)
        return self._collect(self.L.amo_run_slicer(pc, secs.size, secs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                   frac.ctypes.data_as(C.POINTER(C.c_double))))

    def slice_packet(self, chips) -> Frame:
        chips, pc = _f32(chips); f = Frame(); self.L.amo_slice_packet(pc, C.byref(f)); return f

    def format_message(self, f: Frame, first: bool) -> str:
        buf = C.create_string_buffer(200)
        self.L.amo_format_message(C.byref(f), int(first), buf, 200); return buf.value.decode()


def ref_available() -> bool:
    if os.path.exists(REF_SO):
        return True
    if os.path.isdir("/root/reference/lib"):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(REF_SO)


class Ref:
    """The unmodified reference preamble_impl / slicer_impl / modes_crc behind a scheduler stand-in."""

    def __init__(self):
        if not ref_available():
            raise RuntimeError("oracle/_ref/libairmodes_ref.so not built and /root/reference absent")
        L = C.CDLL(REF_SO)
        vp, u64, f32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_float)
        L.aref_run.argtypes = [f32p, f32p, u64, C.c_float, C.c_float, C.c_int]; L.aref_run.restype = vp
        L.aref_run_tagged.argtypes = [f32p, f32p, u64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint64, C.c_double]
        L.aref_run_tagged.restype = vp
        L.aref_run_slicer.argtypes = [f32p, u64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.aref_run_slicer.restype = vp
        L.aref_num_det.argtypes = [vp]; L.aref_num_det.restype = u64
        L.aref_num_calls.argtypes = [vp]; L.aref_num_calls.restype = u64
        L.aref_get_det.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double), f32p]
        L.aref_num_msgs.argtypes = [vp]; L.aref_num_msgs.restype = u64
        L.aref_msg.argtypes = [vp, u64]; L.aref_msg.restype = C.c_char_p
        L.aref_free.argtypes = [vp]
        L.aref_crc.argtypes = [C.c_char_p, C.c_int]; L.aref_crc.restype = C.c_uint32
        L.aref_preamble_params.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.POINTER(C.c_int)]
        self.L = L

    def crc24(self, data: bytes) -> int:
        return int(self.L.aref_crc(bytes(data), len(data)))

    def preamble_params(self, rate, threshold_db):
        r, t, h = C.c_float(), C.c_float(), C.c_int()
        self.L.aref_preamble_params(rate, threshold_db, C.byref(r), C.byref(t), C.byref(h))
        return r.value, t.value, h.value

    def _collect(self, h) -> Result:
        L = self.L
        n = int(L.aref_num_det(h))
        index = np.empty(n, np.uint64); secs = np.empty(n, np.uint64)
        frac = np.empty(n, np.float64); chips = np.empty((n, 240), np.float32)
        L.aref_get_det(h, index.ctypes.data_as(C.POINTER(C.c_uint64)), secs.ctypes.data_as(C.POINTER(C.c_uint64)),
                       frac.ctypes.data_as(C.POINTER(C.c_double)), chips.ctypes.data_as(C.POINTER(C.c_float)))
        msgs = [L.aref_msg(h, k).decode() for k in range(int(L.aref_num_msgs(h)))]
        calls = int(L.aref_num_calls(h))
        L.aref_free(h)
        return Result(index, secs, frac, chips, msgs, None, calls)

    def run_streams(self, bb, avg, rate, threshold_db, slice_=True, start_time=None) -> Result:
        bb, pb = _f32(bb); avg, pa = _f32(avg)
        if start_time is not None:
            return self._collect(self.L.aref_run_tagged(pb, pa, bb.size, rate, threshold_db, int(slice_), 1,
                                                        int(start_time[0]), float(start_time[1])))
        return self._collect(self.L.aref_run(pb, pa, bb.size, rate, threshold_db, int(slice_)))

    def run_slicer(self, chips, secs, frac) -> Result:
        chips, pc = _f32(chips)
        secs = np.ascontiguousarray(secs, np.uint64); frac = np.ascontiguousarray(frac, np.float64)
        return self._collect(self.L.aref_run_slicer(pc, secs.size, secs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    frac.ctypes.data_as(C.POINTER(C.c_double))))
