"""GNU Radio adapter: the B200 receive path as a drop-in for air_modes.rx_path inside a flowgraph.

Importable only where gnuradio is installed (it is not in the build container, so this module is
documented in INTEGRATION.md but not exercised by the tests). radio.py:55-56,73-76 keep working:
`self.connect(source, rx_path(rate, threshold, queue, pmf, dcblock))`.
"""
import numpy as np
from gnuradio import gr  # noqa: F401  (ImportError here is the import guard)

from . import blocks


class rx_path(gr.sync_block):
    def __init__(self, rate, threshold, queue, use_pmf=False, use_dcblock=False):
        gr.sync_block.__init__(self, "modes_rx_path_b200", in_sig=[np.complex64], out_sig=None)
        self._impl = blocks.rx_path(rate, threshold, queue, use_pmf, use_dcblock)
        for name in ("set_rate", "set_threshold", "set_pmf", "get_pmf", "get_threshold"):
            setattr(self, name, getattr(self._impl, name))

    def work(self, input_items, output_items):
        self._impl.process(input_items[0])      # chunking does not change results
        return len(input_items[0])

    def stop(self):
        self._impl.process(np.zeros(0, np.complex64), flush=True)
        return True
