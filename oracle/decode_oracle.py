"""CPU restatement of the reference's per-message decode - SURVEY.md 8 row f4.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else). The product path
(gr_air_modes_b200/csrc/amb_decode.cu) never calls it.

What it restates (file:line into /root/reference/python):
  parse.py:27-87     data_field: fields are bit ranges, 1-based from the MSB, get_bits() -> 0 when the shift is negative
  parse.py:89-231    field tables of bds09_reply, me_reply, tcas_reply, mb_reply, modes_reply; NoHandlerError rules
  parse.py:233-254   decode_id (squawk)
  parse.py:257-273   charmap / parseBDS08 (ident)
  parse.py:276-286   parseBDS05 / parseBDS06 (altitude, ground track, CPR)
  parse.py:288-372   parseBDS09_0/_1/_3, parseBDS62
  parse.py:374-420   MB ident / TCAS subfields
  altitude.py:28-108 decode_alt, gray2bin
  cpr.py:33-62       nz, dlat, nl, dlon
  cpr.py:89-153      cpr_resolve_global (incl. the Python-3 true division inside `zone`, :148)
  cpr.py:158-181     range_bearing
  cpr.py:183-240     cpr_decoder (even/odd lists per ICAO, 10 s / 25 s expiry, newest-first rule)

One deliberate difference, shared with the product: cpr_decoder stamps reports with time.time() at parse time
(cpr.py:219-221); a batch decoder has no wall clock, so `now` is the message's own timestamp secs + frac
(preamble_impl.cc:100-137 via the message text). tests/golden/make_decode_golden.py pins this restatement against
the UNMODIFIED reference modules with time.time patched to exactly that value.

The record layout mirrors struct amb_fields (include/airmodes_b200.h).
"""
from __future__ import annotations

import math

FS_NO_HANDLER, FS_METRIC_ALT, FS_CPR_NO_POS, FS_CPR_STRADDLE = 0x01, 0x02, 0x04, 0x08
FS_HAS_POS, FS_HAS_RANGE, FS_METRIC_THREAT, FS_NOT_QUEUED = 0x10, 0x20, 0x40, 0x80
NO_ALT = -(1 << 31)


def bits(data: int, numbits: int, start: int, num: int) -> int:
    """Message bits start..start+num-1 (1-based, MSB first) of a numbits-bit word (parse.py:71-87, offset 1)."""
    sh = numbits - start - num + 1
    if sh < 0:
        return 0                      # parse.py:83-86: ValueError swallowed, field reads 0
    return (data >> sh) & ((1 << num) - 1)


def gray2bin(g: int) -> int:          # altitude.py:110-117
    i = g >> 1
    while i:
        g ^= i
        i >>= 1
    return g


def decode_alt(alt: int, bit13: bool):
    """altitude.py:28-108. Returns feet or None for MetricAltError."""
    if (alt & 0x40) and bit13:
        return None
    if alt & 0x10:
        if bit13:
            t = ((alt & 0x3F80) >> 2) | ((alt & 0x20) >> 1)
        else:
            t = (alt & 0x1FE0) >> 1
        return ((alt & 0x0F) | t) * 25 - 1000
    if not bit13:
        alt = (alt & 0x3F) | (alt & (0x0FC0 << 1))        # altitude.py:67 as written (operator precedence)
    big = (((alt & 0x0002) >> 1) + ((alt & 0x0008) >> 2) + ((alt & 0x0020) >> 3) + ((alt & 0x0080) >> 4)
           + ((alt & 0x0200) >> 5) + ((alt & 0x0800) >> 6) + ((alt & 0x0001) << 6) + ((alt & 0x0004) << 5))
    d = gray2bin(big)
    c = gray2bin(((alt & 0x0100) >> 8) + ((alt & 0x0400) >> 9) + ((alt & 0x1000) >> 10))
    if c == 7:
        c = 5
    if d % 2:
        c = 6 - c
    return d * 500 + c * 100 - 1300


def decode_id(v: int) -> int:         # parse.py:233-254
    a = ((v & 0x0800) >> 11) + ((v & 0x0200) >> 8) + ((v & 0x0080) >> 5)
    b = ((v & 0x0020) >> 5) + ((v & 0x0008) >> 2) + ((v & 0x0002) << 1)
    c = ((v & 0x1000) >> 12) + ((v & 0x0400) >> 9) + ((v & 0x0100) >> 6)
    d = ((v & 0x0010) >> 2) + ((v & 0x0004) >> 1) + ((v & 0x0001) << 2)
    return a * 1000 + b * 100 + c * 10 + d


def ident48(v: int) -> str:           # parse.py:257-279 / :374-378
    out = ""
    for i in range(8):
        d = (v >> (42 - 6 * i)) & 0x3F
        if 0 < d < 27:
            out += chr(ord("A") + d - 1)
        elif 47 < d < 58:
            out += chr(ord("0") + d - 48)
        else:
            out += " "
    return out


# ---- CPR (cpr.py) -------------------------------------------------------------------------------
def nl(lat: float):
    if abs(lat) >= 87.0:
        return 1.0
    return math.floor((2.0 * math.pi) * math.acos(
        1.0 - (1.0 - math.cos(math.pi / 30.0)) / math.cos((math.pi / 180.0) * abs(lat)) ** 2) ** -1)


def resolve_global(even, odd, mypos, mostrecent: int, surface: int):
    """cpr.py:89-153. Returns ("ok", lat, lon) | ("nopos",) | ("straddle",)."""
    if surface and mypos is None:
        return ("nopos",)
    span = 90.0 if surface else 360.0
    dle, dlo = span / 60, span / 59
    elat, elon, olat, olon = float(even[0]), float(even[1]), float(odd[0]), float(odd[1])
    j = math.floor(((59 * elat - 60 * olat) / 2 ** 17) + 0.5)
    rle = dle * ((j % 60) + elat / 2 ** 17)
    rlo = dlo * ((j % 59) + olat / 2 ** 17)
    if rle > 270.0:
        rle -= 360.0
    if rlo > 270.0:
        rlo -= 360.0
    if nl(rle) != nl(rlo):
        return ("straddle",)
    rlat = rlo if mostrecent else rle
    if surface and mypos[0] < 0:
        rlat -= 90
    n = nl(rlat)
    dl = span / max(n - mostrecent, 1)
    m = math.floor(((elon * (n - 1) - olon * n) / 2 ** 17) + 0.5)
    enclon = olon if mostrecent else elon
    rlon = dl * ((m % max(n - mostrecent, 1)) + enclon / 2. ** 17)
    if surface:
        wat = mypos[1]
        if wat < 0:
            wat += 360
        rlon += 90 * (int(wat) / 90) - 90 * (int(rlon) / 90)       # cpr.py:148: true division under Python 3
    if rlon > 180:
        rlon -= 360.0
    return ("ok", rlat, rlon)


def range_bearing(a, b):              # cpr.py:158-181
    e2 = (1 / 298.257223563) * (2 - (1 / 298.257223563))
    r_mi = 3963.19059 * (math.pi / 180)
    dlat, dlon = b[0] - a[0], b[1] - a[1]
    avg = ((a[0] + b[0]) / 2.0) * math.pi / 180
    r1 = r_mi * (1.0 - e2) / pow((1.0 - e2 * pow(math.sin(avg), 2)), 1.5)
    r2 = r_mi / math.sqrt(1.0 - e2 * pow(math.sin(avg), 2))
    north = r1 * dlat
    east = r2 * math.cos(avg) * dlon
    brg = math.atan2(east, north) * (180.0 / math.pi)
    if brg < 0.0:
        brg += 360.0
    return math.hypot(east, north), brg


class CprState:
    """cpr.py:183-240 with `now` supplied by the caller."""

    def __init__(self, my_location=None):
        self.my_location = my_location
        self.lists = {(0, 0): {}, (1, 0): {}, (0, 1): {}, (1, 1): {}}     # (format, surface) -> icao -> [lat, lon, t]

    def decode(self, icao, lat, lon, fmt, surface, now):
        self.lists[(1 if fmt else 0, surface)][icao] = [lat, lon, now]
        for (f, s), lst in self.lists.items():                            # weed_poslists (cpr.py:196-204)
            lim = 25 if s else 10
            for k in [k for k, it in lst.items() if now - it[2] > lim]:
                del lst[k]
        ev, od = self.lists[(0, surface)], self.lists[(1, surface)]
        if icao not in ev or icao not in od:
            return ("nopos",)
        newer = 1 if (od[icao][2] - ev[icao][2]) > 0 else 0
        return resolve_global(ev[icao][0:2], od[icao][0:2], self.my_location, newer, surface)


def _blank(df, ecc):
    return {"icao": ecc, "ecc": ecc, "df": df, "status": 0, "bds": 0, "subtype": 0xFF, "ca": 0, "fs": 0, "vs": 0,
            "ri": 0, "sl": 0, "cc": 0, "dr": 0, "um": 0, "ftc": 0, "cat": 0, "cpr_format": 0, "surface": 0, "eps": 0,
            "ast": 0, "bds2": 0, "tti": 0, "altitude": NO_ALT, "squawk": -1, "threat_alt": NO_ALT, "cpr_lat": 0,
            "cpr_lon": 0, "aux": [0, 0, 0, 0], "ident": "", "lat": math.nan, "lon": math.nan, "range": math.nan,
            "bearing": math.nan, "val": [math.nan] * 4}


def _alt(r, code, bit13, key="altitude"):
    a = decode_alt(code, bit13)
    if a is None:
        r["status"] |= FS_METRIC_ALT if key == "altitude" else FS_METRIC_THREAT
    else:
        r[key] = a


def decode_one(payload_hex: str, ecc: int, now: float, cpr: CprState) -> dict:
    """One queued message -> record. `payload_hex` is message token 1, `ecc` token 2 (parse.py:425)."""
    data = int(payload_hex, 16)
    nb = 112 if data > (1 << 56) else 56                  # parse.py:222-226
    g = lambda s, n: bits(data, nb, s, n)                 # noqa: E731
    df = g(1, 5)
    r = _blank(df, ecc)
    if df not in (0, 4, 5, 11, 16, 17, 20, 21, 24):       # parse.py:210-220 -> NoHandlerError
        r["status"] |= FS_NO_HANDLER
        return r
    if df in (0, 16):
        r["vs"], r["sl"], r["ri"] = g(6, 1), g(9, 3), g(14, 4)
        if df == 0:
            r["cc"] = g(7, 1)
        _alt(r, g(20, 13), True)
    elif df in (4, 5, 20, 21):
        r["fs"], r["dr"], r["um"] = g(6, 3), g(9, 5), g(14, 6)
        if df in (20, 21):
            mb = g(33, 56)                                  # "mb": (33,56, mb_reply); 0 for a short reply (parse.py:83-86)
            b = lambda s, n: bits(mb, 56, s - 32, n)        # noqa: E731  mb_reply/tcas_reply fields, message coordinates
            bds1, bds2 = b(33, 4), b(37, 4)
            r["bds"], r["bds2"] = bds1, bds2
            if bds1 > 3 or bds2 != 0:                      # parse.py:185-190
                r["status"] |= FS_NO_HANDLER
                return _only_header(r)
            if bds1 == 1:
                r["aux"] = [b(45, 20), b(65, 16), b(81, 8), b(41, 4)]          # acs, bcs, ecs, cfs
            elif bds1 == 2:
                r["ident"] = ident48(b(41, 48))
            elif bds1 == 3:
                tti = b(61, 2)
                r["tti"] = tti
                if tti == 3:                               # tcas_reply has no type 3 (parse.py:157-165)
                    r["status"] |= FS_NO_HANDLER
                    return _only_header(r)
                r["aux"][0], r["aux"][1], r["aux"][2] = b(41, 14), b(55, 4), b(59, 1) | (b(60, 1) << 1)
                if tti == 1:
                    r["aux"][3] = b(63, 26)
                elif tti == 2:
                    r["aux"][3] = b(76, 7) | (b(83, 6) << 8)
                    _alt(r, b(63, 13), True, "threat_alt")                   # parse.py:407
        if df in (4, 20):
            _alt(r, g(20, 13), True)
        else:
            r["squawk"] = decode_id(g(20, 13))
    elif df == 11:
        r["ca"], r["icao"] = g(6, 3), g(9, 24)
    elif df == 24:
        pass                                               # ke/nd/md: no consumer in the reference
    elif df == 17:
        r["ca"], r["icao"] = g(6, 3), g(9, 24)
        me = g(33, 56)
        m = lambda s, n: bits(me, 56, s, n)                # noqa: E731
        ftc = m(1, 5)
        r["ftc"] = ftc
        if 1 <= ftc <= 4:
            r["bds"] = 0x08
            r["cat"], r["ident"] = m(6, 3), ident48(m(9, 48))
        elif 5 <= ftc <= 8:
            r["bds"] = 0x06
            r["surface"], r["cpr_format"], r["cpr_lat"], r["cpr_lon"] = 1, m(22, 1), m(23, 17), m(40, 17)
            r["val"][0] = m(14, 7) * 360. / 128             # parse.py:283
            _cpr(r, cpr, now)
        elif 9 <= ftc <= 18 and ftc != 15:
            r["bds"] = 0x05
            r["cpr_format"], r["cpr_lat"], r["cpr_lon"] = m(22, 1), m(23, 17), m(40, 17)
            _alt(r, m(9, 12), False)                        # parse.py:277 (bit13 False never raises)
            _cpr(r, cpr, now)
        elif ftc == 19:
            r["bds"] = 0x09
            sub = m(6, 3)
            if sub == 0:
                r["subtype"] = 0
                vs = m(42, 9) * 32
                if m(41, 1):
                    vs = 0 - vs
                tr = m(35, 6) * 15 / 62
                if m(34, 1):
                    tr = 0 - tr
                ns, ew = m(23, 11) - 1, m(11, 11) - 1
                vel = math.hypot(ns, ew)
                if m(10, 1):
                    ew = 0 - ew
                if m(22, 1):
                    ns = 0 - ns
                hdg = math.atan2(ew, ns) * (180.0 / math.pi)
                if hdg < 0:
                    hdg += 360
                r["val"] = [vel, hdg, float(vs), float(tr)]
            elif sub in (1, 2):
                r["subtype"] = 1
                geo = m(50, 6) * 25
                if m(49, 1):
                    geo = 0 - geo
                vs = float(m(38, 9) - 1) * 64
                if m(37, 1):
                    vs = 0 - vs
                ns, ew = float(m(26, 10)), float(m(15, 10))
                if sub == 2:
                    ns *= 4
                    ew *= 4
                vel = math.hypot(ns, ew)
                if m(14, 1):
                    ew = 0 - ew
                hdg = 0 if ns == 0 else math.atan(float(ew) / float(ns)) * (180.0 / math.pi)
                if m(25, 1):
                    hdg = 180 - hdg
                if hdg < 0:
                    hdg += 360
                r["val"] = [vel, float(hdg), vs, float(geo)]
            elif sub in (3, 4):
                r["subtype"] = 3
                r["ast"] = m(25, 1)
                vel = m(26, 10)
                if sub == 4:
                    vel *= 4
                vs = float(m(38, 9) - 1) * 64
                if m(37, 1) == 1:
                    vs = 0 - vs
                r["val"] = [m(14, 1) * 360. / 1024, float(vel), vs, float(m(50, 6) - 1) * 25]   # parse.py:356: "mhs"
            else:
                r["status"] |= FS_NO_HANDLER                # bds09_reply.get_type() -> None (parse.py:110-117)
                return _only_header(r, keep_icao=True)
        elif ftc == 28:
            r["bds"] = 0x61
            r["eps"] = m(9, 3)
        else:
            r["status"] |= FS_NO_HANDLER                    # parse.py:140-152
            return _only_header(r, keep_icao=True)
    return r


def _only_header(r, keep_icao=False):
    b = _blank(r["df"], r["ecc"])
    b["status"] = r["status"]
    if keep_icao:
        b["icao"] = r["icao"]
    return b


def _cpr(r, cpr: CprState, now):
    res = cpr.decode(r["icao"], r["cpr_lat"], r["cpr_lon"], r["cpr_format"], r["surface"], now)
    if res[0] == "ok":
        r["status"] |= FS_HAS_POS
        r["lat"], r["lon"] = res[1], res[2]
        if cpr.my_location is not None:
            r["range"], r["bearing"] = range_bearing(cpr.my_location, [res[1], res[2]])
            r["status"] |= FS_HAS_RANGE
    elif res[0] == "straddle":
        r["status"] |= FS_CPR_NO_POS | FS_CPR_STRADDLE
    else:
        r["status"] |= FS_CPR_NO_POS


def decode_batch(msgs, my_location=None, cpr: CprState | None = None):
    """msgs: iterable of (payload_hex, ecc, secs, frac) in stream order -> list of records."""
    cpr = cpr or CprState(my_location)
    return [decode_one(h, e, float(s) + float(f), cpr) for h, e, s, f in msgs]
