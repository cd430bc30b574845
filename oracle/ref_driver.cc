/* TEST INFRASTRUCTURE ONLY (oracle/). Never linked into, imported by or called from the product
 * (gr_air_modes_b200/, include/). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 *
 * What this is: a hand-written "scheduler" that drives the UNMODIFIED reference sources
 *   /root/reference/lib/preamble_impl.cc   (gr::air_modes::preamble_impl::general_work)
 *   /root/reference/lib/slicer_impl.cc     (gr::air_modes::slicer_impl::work, llslicer)
 *   /root/reference/lib/modes_crc.cc       (modes_check_crc)
 * compiled where they lie against the GNU Radio stand-in headers in oracle/shim/. No reference
 * source is copied into this repository; oracle/Makefile compiles them from /root/reference and
 * puts the result in oracle/_ref/ (git-ignored).
 *
 * Canonical semantics ("infinite buffer"): the whole stream is handed to
 * general_work() with all remaining samples on every call; history()-1 zeros are prepended once
 * (GNU Radio pre-fills history with zeros); zero slack follows the end of the stream.
 */
#include <gnuradio/block.h>
#include <gnuradio/msg_queue.h>
#include <gr_air_modes/types.h>
#include <gr_air_modes/modes_crc.h>
#include "preamble_impl.h"
#include "slicer_impl.h"

#include <cstring>
#include <cmath>

namespace {

struct ref_result {
    std::vector<uint64_t> det_index; /* reported sample index = nitems_read + i (preamble_impl.cc:224) */
    std::vector<uint64_t> det_secs;
    std::vector<double> det_frac;
    std::vector<float> chips;        /* 240 per detection (preamble_impl.cc:219-221) */
    std::vector<std::string> msgs;   /* slicer_impl.cc:186-194 */
    uint64_t calls = 0;
};

void run_slicer(ref_result* r, int rate_int)
{
    (void)rate_int;
    gr::msg_queue::sptr q = std::make_shared<gr::msg_queue>();
    gr::air_modes::slicer_impl slicer(q);
    const size_t nchips = r->chips.size();
    std::vector<float> buf(nchips + 1024, 0.0f);
    if (nchips) std::memcpy(buf.data(), r->chips.data(), nchips * sizeof(float));
    for (size_t k = 0; k < r->det_index.size(); k++) {
        gr::tag_t t;
        t.offset = 240ull * k;
        t.key = pmt::string_to_symbol("preamble_found");
        t.value = pmt::make_tuple(pmt::from_uint64(r->det_secs[k]), pmt::from_double(r->det_frac[k]));
        slicer.shim_in_tags.push_back(t);
    }
    slicer.shim_nread = 0;
    gr_vector_const_void_star in(1);
    gr_vector_void_star out;
    in[0] = buf.data();
    /* size = noutput_items - d_check_width(480) must cover every tag offset (slicer_impl.cc:107-114) */
    slicer.work((int)nchips + 480, in, out);
    r->msgs = q->msgs;
}

}  // namespace

extern "C" {

/* Drive the reference preamble + slicer over whole float streams `bb` (stream 0) and `avg`
 * (stream 1) of n items each. */
__attribute__((visibility("default")))
void* aref_run_tagged(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db, int run_slice,
                      int have_tag, uint64_t tag_secs, double tag_frac);

__attribute__((visibility("default")))
void* aref_run(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db, int run_slice)
{
    return aref_run_tagged(bb, avg, n, rate, threshold_db, run_slice, 0, 0, 0.0);
}

/* Same with an rx_time stream tag at item 0 (what a UHD source attaches at stream start; preamble_impl.cc:164-170). */
__attribute__((visibility("default")))
void* aref_run_tagged(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db, int run_slice,
                      int have_tag, uint64_t tag_secs, double tag_frac)
{
    ref_result* r = new ref_result();
    gr::air_modes::preamble_impl blk(rate, threshold_db);
    if (have_tag) {
        gr::tag_t t;
        t.offset = 0;
        t.key = pmt::string_to_symbol("rx_time");
        t.value = pmt::make_tuple(pmt::from_uint64(tag_secs), pmt::from_double(tag_frac));
        blk.shim_in_tags.push_back(t);
    }
    const uint64_t H = blk.history() - 1;
    const uint64_t slack = 64 + (uint64_t)(40.0 * (rate / 2.0e6));
    /* one pad element in front: the (unused) early-gate correlation reads in[i-1] (preamble_impl.cc:187) */
    std::vector<float> s0(1 + H + n + slack, 0.0f), s1(1 + H + n + slack, 0.0f);
    float* a0 = s0.data() + 1;
    float* a1 = s1.data() + 1;
    if (n) {
        std::memcpy(a0 + H, bb, n * sizeof(float));
        std::memcpy(a1 + H, avg, n * sizeof(float));
    }
    const uint64_t ntot = n + H;
    const int rate_int = (int)blk.get_rate(); /* d_sample_rate is int, returned as float */
    uint64_t pos = 0;
    float out[240];
    for (;;) {
        uint64_t remaining = ntot - pos;
        if (remaining > 0x7fffffffull) remaining = 0x7fffffffull; /* int interface; callers keep n < 2^31 */
        gr_vector_int nin(2, (int)remaining);
        gr_vector_const_void_star in(2);
        gr_vector_void_star outs(1);
        in[0] = a0 + pos;
        in[1] = a1 + pos;
        outs[0] = out;
        blk.shim_nread = pos;
        blk.shim_nwritten = r->chips.size();
        blk.shim_consumed = 0;
        blk.shim_out_tags.clear();
        int ret = blk.general_work(240, nin, in, outs);
        r->calls++;
        if (ret == 240) {
            const gr::tag_t& t = blk.shim_out_tags.back();
            uint64_t secs = pmt::to_uint64(pmt::tuple_ref(t.value, 0));
            double frac = pmt::to_double(pmt::tuple_ref(t.value, 1));
            r->det_secs.push_back(secs);
            r->det_frac.push_back(frac);
            r->det_index.push_back(have_tag ? (pos + (uint64_t)blk.shim_consumed - (uint64_t)(240 * (rate / 2000000)))   /* integer spc only */
                                            : secs * (uint64_t)rate_int + (uint64_t)llround(frac * (double)rate_int));
            r->chips.insert(r->chips.end(), out, out + 240);
        }
        pos += (uint64_t)blk.shim_consumed;
        if (blk.shim_consumed == 0 && ret == 0) break;
    }
    if (run_slice) run_slicer(r, rate_int);
    return r;
}

/* Slicer only: chips (240 per detection) + the tag values. */
__attribute__((visibility("default")))
void* aref_run_slicer(const float* chips, uint64_t ndet, const uint64_t* secs, const double* frac)
{
    ref_result* r = new ref_result();
    r->chips.assign(chips, chips + 240 * ndet);
    r->det_secs.assign(secs, secs + ndet);
    r->det_frac.assign(frac, frac + ndet);
    r->det_index.assign(ndet, 0);
    run_slicer(r, 0);
    return r;
}

__attribute__((visibility("default"))) uint64_t aref_num_det(void* h) { return ((ref_result*)h)->det_index.size(); }
__attribute__((visibility("default"))) uint64_t aref_num_calls(void* h) { return ((ref_result*)h)->calls; }
__attribute__((visibility("default")))
void aref_get_det(void* h, uint64_t* idx, uint64_t* secs, double* frac, float* chips)
{
    ref_result* r = (ref_result*)h;
    size_t n = r->det_index.size();
    if (idx) std::memcpy(idx, r->det_index.data(), n * sizeof(uint64_t));
    if (secs) std::memcpy(secs, r->det_secs.data(), n * sizeof(uint64_t));
    if (frac) std::memcpy(frac, r->det_frac.data(), n * sizeof(double));
    if (chips) std::memcpy(chips, r->chips.data(), n * 240 * sizeof(float));
}
__attribute__((visibility("default"))) uint64_t aref_num_msgs(void* h) { return ((ref_result*)h)->msgs.size(); }
__attribute__((visibility("default"))) const char* aref_msg(void* h, uint64_t k) { return ((ref_result*)h)->msgs[k].c_str(); }
__attribute__((visibility("default"))) void aref_free(void* h) { delete (ref_result*)h; }

/* Reference CRC (modes_crc.cc:55-63). */
__attribute__((visibility("default")))
uint32_t aref_crc(const uint8_t* data, int length)
{
    std::vector<unsigned char> tmp(data, data + length);
    return (uint32_t)modes_check_crc(tmp.data(), length);
}

/* Accessors of the reference block (preamble_impl.cc:56-76). */
__attribute__((visibility("default")))
void aref_preamble_params(float rate, float threshold_db, float* get_rate, float* get_threshold, int* history)
{
    gr::air_modes::preamble_impl blk(rate, threshold_db);
    *get_rate = blk.get_rate();
    *get_threshold = blk.get_threshold();
    *history = (int)blk.history();
}

}  // extern "C"
