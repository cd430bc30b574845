"""SURVEY.md 8 row f4 (field decode: parse.py / altitude.py / cpr.py) - CPU side.

  * oracle/decode_oracle.py against the committed golden produced by the UNMODIFIED reference modules
    (tests/golden/make_decode_golden.py), bit-exact incl. every float;
  * the same against the live reference where /root/reference exists;
  * the product's per-message device code (gr_air_modes_b200/csrc/amb_decode_core.h, the `__host__ __device__`
    functions the CUDA kernels call) compiled for the host by tests/decode_host_shim.cc - a test-only harness, not a
    product path - against the golden, so the decode arithmetic is checked even where no GPU is present.
"""
import os
import sys

import pytest

from helpers import compare_decode, load_decode_golden

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_decode_oracle_matches_reference_golden():
    from oracle import decode_oracle as do
    n = 0
    for ci, case in enumerate(load_decode_golden()):
        recs = do.decode_batch([tuple(m) for m in case["msgs"]], case["location"])
        assert len(recs) == len(case["ref"])
        for k, (rec, ref) in enumerate(zip(recs, case["ref"])):
            compare_decode(rec, ref, 0.0, "case %d msg %d" % (ci, k))
            n += 1
    assert n > 5000


@pytest.mark.skipif(not os.path.isdir("/root/reference/python"), reason="reference tree not present")
def test_decode_oracle_matches_live_reference_on_fresh_seeds():
    import decode_cases
    from oracle import decode_oracle as do
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_decode_golden as mg
    mods = mg.load_reference()
    try:
        for seed in (101, 102, 103):
            loc, msgs = decode_cases.make_case(seed, location=(None if seed == 102 else (40.0 + seed % 7, -70.0)),
                                               seconds=30.0, surface_share=0.4)
            ref = mg.run_reference(loc, msgs, mods)
            recs = do.decode_batch(msgs, loc)
            for k, (rec, r) in enumerate(zip(recs, ref)):
                compare_decode(rec, r, 0.0, "seed %d msg %d" % (seed, k))
    finally:
        mg.unload_reference()
