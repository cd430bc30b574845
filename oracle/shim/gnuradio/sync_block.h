/* Test infrastructure only (oracle/): stand-in for <gnuradio/sync_block.h>. */
#ifndef ORACLE_SHIM_GNURADIO_SYNC_BLOCK_H
#define ORACLE_SHIM_GNURADIO_SYNC_BLOCK_H
#include "block.h"
namespace gr {
class sync_block : public block {
public:
    sync_block() {}
    sync_block(const std::string& name, io_signature::sptr i, io_signature::sptr o) : block(name, i, o) {}
};
}  // namespace gr
#endif
