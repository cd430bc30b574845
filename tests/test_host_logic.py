"""CPU-side checks of the product's host logic and of the C-ABI surface (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gr_air_modes_b200 as am
from gr_air_modes_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_native()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "airmodes_b200.h")).read()
    return sorted(set(re.findall(r"AMB_API\s+[\w\s\*]+?\b(amb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 25
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert set(names) == bound
    for n in names:
        assert getattr(lib, n) is not None


def test_frame_layout_matches_header():
    assert C.sizeof(_lib.Frame) == 80
    assert _lib.Frame.data.offset == 60 and _lib.Frame.lowconfbits.offset == 36


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.amb_create(0, 4e6, 7.0, 1, 0, C.byref(h)) == -2      # AMB_ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        am.rx_path(4e6, 7.0, am.msg_queue(), use_pmf=True)
    with pytest.raises(RuntimeError):
        am.preamble(4e6, 7.0)


def test_host_crc_known_answers(lib):
    assert am.modes_check_crc(bytes.fromhex("8D4840D6202CC371C32CE0576098"), 11) == 0x576098
    assert am.modes_check_crc(bytes.fromhex("8D40621D58C382D690C8AC2863A7"), 11) == 0x2863A7
    assert am.modes_check_crc(bytes([0, 0, 1]), 3) == 0xFFF409


def test_host_crc_matches_oracle(lib, port):
    rng = np.random.default_rng(3)
    for length in (4, 11):
        for _ in range(300):
            b = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
            assert am.modes_check_crc(b, length) == port.crc24(b)


def test_message_format_matches_oracle(lib, port):
    from oracle import cpu_oracle as co
    rng = np.random.default_rng(4)
    for k in range(200):
        f = _lib.Frame()
        g = co.Frame()
        nb = 112 if k % 2 else 56
        data = rng.integers(0, 256, 14, dtype=np.uint8)
        for m in range(14):
            f.data[m] = g.data[m] = int(data[m])
        f.nbits = g.nbits = nb
        f.crc = g.crc = int(rng.integers(0, 1 << 24))
        ref = np.float32(10.0 ** rng.uniform(-6, 1))
        f.ref_level = g.ref_level = float(ref)
        f.secs = g.secs = int(rng.integers(0, 1000))
        f.frac = g.frac = float(rng.random())
        for first in (True, False):
            assert am.format_message(f, first) == port.format_message(g, first)


def test_msg_queue_stand_in():
    q = am.msg_queue()
    assert q.empty_p()
    q.handle(am.message_from_string("abc"))
    q.insert_tail(am.message("def"))
    assert q.count() == 2 and q.delete_head().to_string() == "abc"
    assert q.strings() == ["def"]
    q.flush()
    assert q.empty_p() and q.delete_head_nowait() is None


def test_product_does_not_touch_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "gr_air_modes_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert all(w not in txt for w in ("cpu_oracle", "liboracle", "modes_oracle", "decode_oracle")), fn
    assert "oracle" not in open(os.path.join(ROOT, "include", "airmodes_b200.h")).read().lower()
    # the user-facing tools and the C example are product too; diagnostics that use the checker live in tests/tools
    for d in ("tools", "examples"):
        for fn in os.listdir(os.path.join(ROOT, d)):
            if fn.endswith((".py", ".sh", ".c")):
                txt = open(os.path.join(ROOT, d, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, fn


def test_geometry_matches_reference_blocks(lib, port, ref):
    """Host-side parameter derivation (no GPU): same numbers as preamble_impl's set_rate/set_threshold and the
    loop bounds its general_work evaluates, for integer and fractional samples/chip."""
    g = _lib.Geometry()
    for rate in (2e6, 2.4e6, 3e6, 3.2e6, 4e6, 5e6, 6e6, 8e6, 10e6, 12.5e6, 16e6, 20e6):
        for thr in (3.0, 7.0, 11.5):
            assert lib.amb_query_geometry(rate, thr, 1, C.byref(g)) == 0
            p = port.params(rate, thr)
            r_rate, r_thr, r_hist = ref.preamble_params(rate, thr)
            assert (float(g.rate_int), g.history) == (r_rate, r_hist)
            assert (g.samples_per_chip, g.samples_per_symbol, g.threshold, g.check_width) == (p.spc, p.sps, p.threshold, p.check_width)
            assert list(g.pulse_offset) == list(p.po)
            sps = np.float32(p.sps)
            qa = [j for j in range(int(1.5 * float(sps)), 4000) if np.float32(j) <= np.float32(3) * sps]
            qb = [j for j in range(int(np.float32(5) * sps), 4000) if float(j) <= 7.5 * float(sps)]
            assert [g.quiet_a[0], g.quiet_a[1]] == [qa[0], qa[-1]] and [g.quiet_b[0], g.quiet_b[1]] == [qb[0], qb[-1]]
            assert g.packet_skip == int(np.float32(240) * np.float32(p.spc)) and g.max_late == max(1, int(np.ceil(p.spc)))
            assert g.pmf_len == int(rate / 2e6) and g.floor_len == 48 * int(rate / 2e6)
    for bad in (1e6, 1.99e6, 21e6, 40e6):
        assert lib.amb_query_geometry(bad, 7.0, 1, C.byref(g)) == -4          # AMB_ERR_RATE


def test_zmq_dl_data_publisher():
    """f3: messages leave as ZMQ multipart [b'dl_data', text], the framing radio.py/zmq_socket.py use."""
    zmq = pytest.importorskip("zmq")
    import time
    from gr_air_modes_b200.zmq_pub import zmq_queue
    ctx = zmq.Context()
    q = zmq_queue("inproc://amb-test-pub", context=ctx)
    sub = ctx.socket(zmq.SUB)
    sub.connect("inproc://amb-test-pub")
    sub.setsockopt(zmq.SUBSCRIBE, b"dl_data")
    time.sleep(0.05)
    texts = ["8d4840d6202cc371c32ce0576098 000000 0.2475689948 5 1.25e-06", "5d4840d6a1b2c3 000000 0.01 6 0.5"]
    for t in texts:
        q.handle(am.message_from_string(t))
    got = []
    for _ in texts:
        assert sub.poll(1000)
        got.append(sub.recv_multipart())
    assert got == [[b"dl_data", t.encode()] for t in texts] and q.sent == 2
    sub.close(linger=0); q.close(); ctx.term()


def test_c_example_builds_against_the_header(lib):
    """A C99 translation unit that only includes airmodes_b200.h links against the library."""
    exe = build.build_c_example(force=True)
    assert os.path.exists(exe)
    import subprocess
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr


def test_fields_layout_matches_header():
    assert C.sizeof(_lib.Fields) == 144
    assert _lib.Fields.altitude.offset == 28 and _lib.Fields.ident.offset == 64 and _lib.Fields.lat.offset == 72
    assert _lib.Fields.val.offset == 104


def test_decoder_has_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gr_air_modes_b200 import decode
    h = C.c_void_p()
    assert lib.amb_decoder_create(0, 0, 0.0, 0.0, C.byref(h)) == -2      # AMB_ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        decode.batch_decoder([37.0, -122.0])


def test_frame_bits_host_helper_follows_get_bits(lib):
    """amb_frame_bits = data_field.get_bits (parse.py:71-87) incl. the 'negative shift reads 0' rule."""
    from gr_air_modes_b200 import decode
    from oracle import decode_oracle as do
    rng = np.random.default_rng(8)
    for k in range(300):
        nbytes = 14 if k % 2 else 7
        raw = bytearray(rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes())
        if nbytes == 14:
            raw[0] |= 0x80                      # long replies have the top DF bit set (slicer_impl.cc:140)
        else:
            raw[0] &= 0x7F
        arr, _ = decode.frames_from_messages([(bytes(raw).hex(), 0, 0, 0.0)])
        v = int.from_bytes(raw, "big")
        for _ in range(20):
            s, n = int(rng.integers(1, 113)), int(rng.integers(1, 57))
            assert decode.frame_bits(arr[0], s, n) == do.bits(v, 8 * nbytes, s, n), (raw.hex(), s, n)


def test_batch_message_formatting_equals_per_frame(lib, port):
    """amb_format_messages + slicer.emit (one library call per batch) give exactly the per-frame amb_format_message
    strings, incl. the stream's first message having precision 6 only once (slicer_impl.cc:192), frames with
    passed == 0 skipped, and the oracle's text."""
    from oracle import cpu_oracle as co
    from gr_air_modes_b200 import blocks
    rng = np.random.default_rng(12)
    for n, p_pass in ((0, 1.0), (1, 1.0), (5, 0.0), (700, 0.8), (3, 0.5)):
        frames = (_lib.Frame * max(n, 1))()
        oracle_frames = []
        for k in range(n):
            f = frames[k]
            g = co.Frame()
            f.nbits = g.nbits = 112 if rng.random() < 0.5 else 56
            data = rng.integers(0, 256, 14, dtype=np.uint8)
            for m in range(14):
                f.data[m] = g.data[m] = int(data[m])
            f.crc = g.crc = int(rng.integers(0, 1 << 24))
            f.ref_level = g.ref_level = float(np.float32(10.0 ** rng.uniform(-30, 8)))
            f.secs = g.secs = int(rng.integers(0, 1 << 40))
            f.frac = g.frac = float(rng.random())
            f.passed = int(rng.random() < p_pass)
            oracle_frames.append(g)
        for first in (True, False):
            want, fst = [], first
            for k in range(n):
                if frames[k].passed:
                    want.append(blocks.format_message(frames[k], fst))
                    assert want[-1] == port.format_message(oracle_frames[k], fst)
                    fst = False
            for as_list in (False, True):
                q = am.msg_queue()
                sl = blocks.slicer.__new__(blocks.slicer)
                sl._queue, sl._first = q, first
                k = sl.emit(list(frames)[:n] if as_list else frames, n)
                assert k == len(want) and q.strings() == want
                assert sl._first == (first and not want)
    # buffer too small -> error, not truncation
    one = (_lib.Frame * 1)()
    one[0].nbits, one[0].passed = 112, 1
    buf = C.create_string_buffer(8)
    assert lib.amb_format_messages(C.cast(one, C.c_void_p), 1, 1, buf, 8) < 0
