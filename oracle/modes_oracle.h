/* TEST INFRASTRUCTURE ONLY (oracle/): plain-C CPU restatement of the Mode S receive hot path of
 * gr-air-modes, used as the parity checker for the CUDA path. The product never includes, links
 * or calls this. Every function cites the reference file:line (relative to /root/reference) it
 * follows. Pinned against the unmodified reference (oracle/_ref, built by oracle/Makefile) by
 * tests/test_oracle_vs_ref.py and by the committed fixtures in tests/golden/.
 */
#ifndef MODES_ORACLE_H
#define MODES_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* moving-average schedules */
#define AMO_MA_CANONICAL 0 /* fp64 ascending window sum, rounded once to fp32, then * scale (fp32) */
#define AMO_MA_GR_FLOAT 1  /* GNU Radio 3.8 moving_average_ff: fp32 running sum restarted every `chunk` outputs */
#define AMO_MA_SLIDING64 2 /* fp64 sliding add/subtract (fast; equals CANONICAL whenever the fp64 sums are exact) */

typedef struct {
    float spc;         /* d_samples_per_chip   (preamble_impl.cc:57) */
    float sps;         /* d_samples_per_symbol (preamble_impl.cc:58) */
    int check_width;   /* preamble_impl.cc:59 */
    int rate_int;      /* d_sample_rate is an int (preamble_impl.h:24, preamble_impl.cc:60) */
    int history;       /* set_history(d_samples_per_symbol) (preamble_impl.cc:62) */
    float threshold_db;
    float threshold;   /* powf(10., dB/20.) (preamble_impl.cc:67) */
    int po[4];         /* pulse_offsets (preamble_impl.cc:158-162) */
} amo_params;

typedef struct {
    uint64_t index;    /* reported sample index nitems_read+i (preamble_impl.cc:224): true index + history-1 */
    uint64_t secs;     /* tag_to_timestamp (preamble_impl.cc:100-137) */
    double frac;
    float ref_level;   /* slicer_impl.cc:128-131 */
    uint32_t crc;      /* crc ^ ap (slicer_impl.cc:173-177) */
    uint8_t nbits;     /* 56 / 112 (slicer_impl.cc:140-142) */
    uint8_t df;        /* message_type (slicer_impl.cc:168) */
    uint8_t numlowconf;
    uint8_t passed;    /* 1 = survives every `continue` in slicer_impl.cc:162-182 and is sent to the queue */
    uint8_t lowconfbits[24];
    uint8_t data[14];
} amo_frame;

void amo_make_params(float channel_rate, float threshold_db, amo_params* p);

/* modes_crc.cc:38-63 */
uint32_t amo_crc24(const uint8_t* data, int length);

/* GNU Radio front end wired in python/rx_path.py:38-65 (a1, a3, a4 of SURVEY.md 8a) */
void amo_mag2(const float* iq, uint64_t n, float* m2);
void amo_moving_average(const float* u, uint64_t n, int length, float scale, int mode, int chunk, float* out);
void amo_frontend(const float* iq, uint64_t n, float rate, int use_pmf, int ma_mode, int chunk,
                  float* bb, float* avg);

/* filter.dc_blocker_cc(D, long_form=False) (rx_path.py:39-41); GNU Radio code, parity unpinned (see .c) */
void amo_dc_blocker(const float* iq, uint64_t n, int D, int mode, float* out);

/* rx_time tag at item 0 used by tag_to_timestamp (preamble_impl.cc:104-116); (0, 0.0) = no tag. Test hook, global. */
void amo_set_start_time(uint64_t secs, double frac);

/* Opaque result of a run. */
typedef struct amo_result amo_result;

/* preamble_impl::general_work driven with infinite-buffer semantics over whole streams. */
amo_result* amo_scan(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db);
/* ... followed by slicer_impl::work on every detection. */
amo_result* amo_run_streams(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db);
/* front end + scan + slice from interleaved IQ. */
amo_result* amo_run_iq(const float* iq, uint64_t n, float rate, float threshold_db, int use_pmf,
                       int ma_mode, int chunk);
amo_result* amo_run_iq_dc(const float* iq, uint64_t n, float rate, float threshold_db, int use_pmf, int use_dcblock,
                          int ma_mode, int chunk);
/* slicer only on 240-chip packets. */
amo_result* amo_run_slicer(const float* chips, uint64_t ndet, const uint64_t* secs, const double* frac);

uint64_t amo_num_det(const amo_result* r);
uint64_t amo_num_calls(const amo_result* r);
void amo_get_det(const amo_result* r, uint64_t* index, uint64_t* secs, double* frac, float* chips);
void amo_get_frames(const amo_result* r, amo_frame* frames); /* one per detection, see .passed */
uint64_t amo_num_msgs(const amo_result* r);
const char* amo_msg(const amo_result* r, uint64_t k);
void amo_free(amo_result* r);

/* llslicer + slicer_impl::work for ONE packet of 240 chips (slicer_impl.cc:67-182). */
void amo_slice_packet(const float* chips, amo_frame* f);
/* slicer_impl.cc:186-192; `first` selects the 6-digit default precision of the first message of a
 * slicer instance (setprecision(10) at :192 is sticky afterwards). Returns strlen. */
int amo_format_message(const amo_frame* f, int first, char* buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif
