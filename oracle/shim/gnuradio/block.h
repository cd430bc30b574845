/* Test infrastructure only (oracle/): stand-in for <gnuradio/block.h>. It records what the
 * block asks of the scheduler (consume_each, tags, history, output multiple) so that a
 * hand-written driver (oracle/ref_driver.cc) can play the scheduler's role. */
#ifndef ORACLE_SHIM_GNURADIO_BLOCK_H
#define ORACLE_SHIM_GNURADIO_BLOCK_H
#include <math.h>
#include <stdint.h>
#include <algorithm>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include "attributes.h"
#include "io_signature.h"
#include "tags.h"

namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T* p) { return std::shared_ptr<T>(p); }
}
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gr {
class block {
public:
    block() {}
    block(const std::string& name, io_signature::sptr, io_signature::sptr) : shim_name(name) {}
    virtual ~block() {}
    std::string name() const { return shim_name; }
    long unique_id() const { return 1; }
    void set_output_multiple(int m) { shim_output_multiple = m; }
    void set_history(unsigned h) { shim_history = h; }
    unsigned history() const { return shim_history; }
    uint64_t nitems_read(unsigned) const { return shim_nread; }
    uint64_t nitems_written(unsigned) const { return shim_nwritten; }
    void consume_each(int n) { shim_consumed = n; }
    void get_tags_in_range(std::vector<tag_t>& v, unsigned, uint64_t start, uint64_t end,
                           const pmt::pmt_t& key) {
        v.clear();
        for (size_t k = 0; k < shim_in_tags.size(); k++) {
            const tag_t& t = shim_in_tags[k];
            if (t.offset >= start && t.offset < end && t.key && key && t.key->sym == key->sym)
                v.push_back(t);
        }
    }
    void add_item_tag(unsigned, uint64_t offset, const pmt::pmt_t& key, const pmt::pmt_t& value,
                      const pmt::pmt_t& srcid) {
        tag_t t; t.offset = offset; t.key = key; t.value = value; t.srcid = srcid;
        shim_out_tags.push_back(t);
    }
    /* driver-visible state */
    std::string shim_name;
    int shim_output_multiple = 1;
    unsigned shim_history = 1;
    uint64_t shim_nread = 0, shim_nwritten = 0;
    int shim_consumed = 0;
    std::vector<tag_t> shim_in_tags, shim_out_tags;
};
}  // namespace gr
#endif
