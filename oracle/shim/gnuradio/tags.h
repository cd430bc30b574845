/* Test infrastructure only (oracle/): stand-in for <gnuradio/tags.h>. */
#ifndef ORACLE_SHIM_GNURADIO_TAGS_H
#define ORACLE_SHIM_GNURADIO_TAGS_H
#include "pmt_shim.h"
namespace gr {
struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key;    /* null by default => tag_to_timestamp takes the "no rx_time" branch */
    pmt::pmt_t value;
    pmt::pmt_t srcid;
};
}  // namespace gr
#endif
