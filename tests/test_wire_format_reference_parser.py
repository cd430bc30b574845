"""the message strings we emit are consumed by the reference's UNMODIFIED parser
(python/parse.py make_parser, :422-436). Runs only where /root/reference exists (this container); the reference
files are imported in place behind a stub `air_modes` package, never copied."""
import importlib.util
import math
import os
import sys
import types

import pytest

from helpers import load_golden

REF = "/root/reference/python"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def ref_parser():
    saved = {k: v for k, v in sys.modules.items() if k == "air_modes" or k.startswith("air_modes.")}
    pkg = types.ModuleType("air_modes")
    pkg.__path__ = []
    sys.modules["air_modes"] = pkg
    try:
        _load("air_modes.exceptions", os.path.join(REF, "exceptions.py"))
        _load("air_modes.altitude", os.path.join(REF, "altitude.py"))
        mt = _load("air_modes.modes_types", os.path.join(REF, "modes_types.py"))
        pkg.modes_report, pkg.stamp = mt.modes_report, mt.stamp
        yield _load("air_modes.parse", os.path.join(REF, "parse.py"))
    finally:
        for k in [k for k in sys.modules if k == "air_modes" or k.startswith("air_modes.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_parser_accepts_our_messages(ref_parser):
    meta, scenes = load_golden()
    seen = 0
    for s, _ in scenes:
        pub = {}
        publish = ref_parser.make_parser(pub)
        for m in s["msgs"]:
            pub.clear()
            publish(m)
            data, ecc, ref, secs, frac = m.split()
            if "modes_dl" not in pub:       # the parser drops types it has no handler for (ADSBError)
                continue
            rep = pub["modes_dl"]
            assert rep.ecc == int(ecc, 16)
            assert abs(rep.rssi - 10.0 * math.log10(max(1e-8, float(ref)))) < 1e-12
            assert rep.timestamp.secs == int(secs) and rep.timestamp.frac_secs == float(frac)
            df = int(data[:2], 16) >> 3
            assert rep.data.get_type() == df and ("type%i_dl" % df) in pub
            if df in (11, 17) and data in s["sent"]:
                assert rep.data["aa"] == int(data[2:8], 16)      # ICAO address straight from our payload
                seen += 1
    assert seen >= 20
