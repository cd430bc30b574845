"""TEST STAND-IN for the `gnuradio` package (not installed in this image): just enough of gnuradio.gr for
gr_air_modes_b200/gr_adapter.py to be imported and driven by tests/test_library_simt.py. Not GNU Radio."""
from . import gr  # noqa: F401
