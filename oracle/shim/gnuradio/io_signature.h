/* Test infrastructure only (oracle/): stand-in for <gnuradio/io_signature.h>. */
#ifndef ORACLE_SHIM_GNURADIO_IO_SIGNATURE_H
#define ORACLE_SHIM_GNURADIO_IO_SIGNATURE_H
#include <memory>
namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    int min_streams, max_streams, size0, size1;
    static sptr make(int mn, int mx, int sz) {
        sptr p = std::make_shared<io_signature>(); p->min_streams = mn; p->max_streams = mx;
        p->size0 = p->size1 = sz; return p;
    }
    static sptr make2(int mn, int mx, int sz0, int sz1) {
        sptr p = make(mn, mx, sz0); p->size1 = sz1; return p;
    }
};
}  // namespace gr
#endif
