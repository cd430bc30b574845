/* TEST INFRASTRUCTURE ONLY (oracle/). See modes_oracle.h. Plain-C restatement of the reference
 * algorithm; citations are file:line under /root/reference. Compile with -ffp-contract=off so
 * that no multiply-add is fused (the reference's x86-64 build does not fuse either).
 *
 * Parity status: PINNED - tests/test_oracle_vs_ref.py checks every function here against the
 * unmodified reference objects in oracle/_ref on seeded scenes (2/4/10/20 Msps and fractional
 * rates), and tests/golden/ holds reference-generated fixtures that travel to the GPU box.
 * The GNU Radio 3.8 front-end blocks (complex_to_mag_squared, moving_average_ff) are NOT in
 * /root/reference; their arithmetic is restated from the published GNU Radio 3.8 algorithm
 * (gr-blocks/lib/moving_average_impl.cc, volk_32fc_magnitude_squared_32f generic kernel) and
 * anchored on the call sites python/rx_path.py:38-65.
 */
#define _POSIX_C_SOURCE 200809L   /* strdup */
#include "modes_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ parameters */

/* preamble_impl.cc:46,56-68 */
void amo_make_params(float channel_rate, float threshold_db, amo_params* p)
{
    const int chip_rate = 2000000;              /* :46 */
    p->spc = channel_rate / chip_rate;          /* :57  float / int */
    p->sps = p->spc * 2;                        /* :58 */
    p->check_width = 120 * p->sps;              /* :59  float -> int */
    p->rate_int = channel_rate;                 /* :60  float -> int */
    p->history = (int)(unsigned)p->sps;         /* :62  set_history(unsigned) */
    p->threshold_db = threshold_db;             /* :66 */
    p->threshold = powf(10., threshold_db / 20.); /* :67 */
    p->po[0] = 0;                               /* :158-162 */
    p->po[1] = (int)(2 * p->spc);
    p->po[2] = (int)(7 * p->spc);
    p->po[3] = (int)(9 * p->spc);
}

/* ------------------------------------------------------------------ CRC */

/* modes_crc.cc:33-52 (table), :55-63 (check). POLY 0xFFF409, init 0, MSB first. */
static uint32_t crc_tab[256];
static int crc_tab_ready = 0;
static void crc_build(void)
{
    for (int n = 0; n < 256; n++) {
        uint32_t crc = (uint32_t)n << 16;
        for (int k = 0; k < 8; k++) {
            if (crc & 0x800000u) crc = ((crc << 1) ^ 0xFFF409u) & 0xFFFFFFu;
            else crc = (crc << 1) & 0xFFFFFFu;
        }
        crc_tab[n] = crc & 0xFFFFFFu;
    }
    crc_tab_ready = 1;
}
uint32_t amo_crc24(const uint8_t* data, int length)
{
    if (!crc_tab_ready) crc_build();
    uint32_t crc = 0;
    for (int i = 0; i < length; i++)
        crc = crc_tab[((crc >> 16) ^ data[i]) & 0xff] ^ (crc << 8);
    return crc & 0xFFFFFFu;
}

/* ------------------------------------------------------------------ front end */

/* blocks.complex_to_mag_squared (rx_path.py:38): re*re + im*im, two roundings + one add, no FMA. */
void amo_mag2(const float* iq, uint64_t n, float* m2)
{
    for (uint64_t k = 0; k < n; k++) {
        float re = iq[2 * k], im = iq[2 * k + 1];
        float a = re * re;
        float b = im * im;
        m2[k] = a + b;
    }
}

/* blocks.moving_average_ff(length, scale) (rx_path.py:49,54). Output n covers inputs n-length+1..n,
 * zeros before the start of the stream (GNU Radio history pre-fill). */
void amo_moving_average(const float* u, uint64_t n, int length, float scale, int mode, int chunk, float* out)
{
    const int64_t L = length;
    if (mode == AMO_MA_CANONICAL) {
        for (uint64_t k = 0; k < n; k++) {
            int64_t lo = (int64_t)k - L + 1;
            if (lo < 0) lo = 0;
            double acc = 0.0;
            for (int64_t j = lo; j <= (int64_t)k; j++) acc += (double)u[j];
            float s = (float)acc;
            out[k] = s * scale;
        }
    } else if (mode == AMO_MA_SLIDING64) {
        double acc = 0.0;
        for (uint64_t k = 0; k < n; k++) {
            acc += (double)u[k];
            if ((int64_t)k >= L) acc -= (double)u[k - L];
            float s = (float)acc;
            out[k] = s * scale;
        }
    } else { /* AMO_MA_GR_FLOAT: fp32 running sum, restarted at each work() call of <= chunk outputs */
        if (chunk <= 0) chunk = 4096;
        for (uint64_t start = 0; start < n; start += (uint64_t)chunk) {
            uint64_t num = n - start < (uint64_t)chunk ? n - start : (uint64_t)chunk;
            float sum = 0.0f;
            for (int64_t i = 0; i < L - 1; i++) {
                int64_t idx = (int64_t)start - (L - 1) + i;
                sum += idx >= 0 ? u[idx] : 0.0f;
            }
            for (uint64_t i = 0; i < num; i++) {
                sum += u[start + i];
                out[start + i] = sum * scale;
                int64_t idx = (int64_t)(start + i) - (L - 1);
                sum -= idx >= 0 ? u[idx] : 0.0f;
            }
        }
    }
}

/* filter.dc_blocker_cc(D, False) (rx_path.py:39-41, D = 100*spc). GNU Radio 3.8 gr-filter/lib/dc_blocker_cc_impl.cc,
 * long_form == false:  y1 = ma0.filter(x); y2 = ma1.filter(y1); out = ma0.delayed_sig() - y2, where
 * moving_averager_c(D).filter keeps a D-sample running complex sum and returns sum / (float)D, and delayed_sig()
 * is the input D-1 samples ago. NOT in /root/reference: restated from the published algorithm; parity UNPINNED.
 *   mode AMO_MA_CANONICAL: each window sum in fp64, ascending, per component, rounded once to fp32, then / (float)D
 *   mode AMO_MA_GR_FLOAT : GNU Radio's recursive fp32 running sum  y = x - x[n-D] + y_prev  (never restarted) */
void amo_dc_blocker(const float* iq, uint64_t n, int D, int mode, float* out)
{
    float* ma0 = (float*)calloc(2 * (n ? n : 1), sizeof(float));
    const float fD = (float)D;
    if (mode == AMO_MA_GR_FLOAT) {
        float s0r = 0, s0i = 0, s1r = 0, s1i = 0;
        for (uint64_t k = 0; k < n; k++) {
            float xr = iq[2 * k], xi = iq[2 * k + 1];
            float or_ = k >= (uint64_t)D ? iq[2 * (k - D)] : 0.0f, oi = k >= (uint64_t)D ? iq[2 * (k - D) + 1] : 0.0f;
            s0r = xr - or_ + s0r; s0i = xi - oi + s0i;
            ma0[2 * k] = s0r / fD; ma0[2 * k + 1] = s0i / fD;
            float pr = k >= (uint64_t)D ? ma0[2 * (k - D)] : 0.0f, pi = k >= (uint64_t)D ? ma0[2 * (k - D) + 1] : 0.0f;
            s1r = ma0[2 * k] - pr + s1r; s1i = ma0[2 * k + 1] - pi + s1i;
            float dr = k >= (uint64_t)(D - 1) ? iq[2 * (k - D + 1)] : 0.0f, di = k >= (uint64_t)(D - 1) ? iq[2 * (k - D + 1) + 1] : 0.0f;
            out[2 * k] = dr - s1r / fD; out[2 * k + 1] = di - s1i / fD;
        }
    } else {
        for (uint64_t k = 0; k < n; k++) {
            int64_t lo = (int64_t)k - D + 1; if (lo < 0) lo = 0;
            double ar = 0.0, ai = 0.0;
            for (int64_t j = lo; j <= (int64_t)k; j++) { ar += (double)iq[2 * j]; ai += (double)iq[2 * j + 1]; }
            ma0[2 * k] = (float)ar / fD; ma0[2 * k + 1] = (float)ai / fD;
        }
        for (uint64_t k = 0; k < n; k++) {
            int64_t lo = (int64_t)k - D + 1; if (lo < 0) lo = 0;
            double ar = 0.0, ai = 0.0;
            for (int64_t j = lo; j <= (int64_t)k; j++) { ar += (double)ma0[2 * j]; ai += (double)ma0[2 * j + 1]; }
            float m1r = (float)ar / fD, m1i = (float)ai / fD;
            float dr = k >= (uint64_t)(D - 1) ? iq[2 * (k - D + 1)] : 0.0f, di = k >= (uint64_t)(D - 1) ? iq[2 * (k - D + 1) + 1] : 0.0f;
            out[2 * k] = dr - m1r; out[2 * k + 1] = di - m1i;
        }
    }
    free(ma0);
}

/* rx_path.py:34-35 (_spc = int(rate/2e6)), :38 (demod), :48-51 (pmf), :54 (floor), :63-64 (wiring). */
void amo_frontend(const float* iq, uint64_t n, float rate, int use_pmf, int ma_mode, int chunk,
                  float* bb, float* avg)
{
    const int spc = (int)((double)rate / 2e6); /* rx_path.py:35: int(rate/2e6) */
    float* m2 = (float*)malloc((n ? n : 1) * sizeof(float));
    amo_mag2(iq, n, m2);
    if (use_pmf) {
        float scale = (float)(1.0 / spc);                  /* rx_path.py:49, double -> float param */
        amo_moving_average(m2, n, spc, scale, ma_mode, chunk, bb);
    } else {
        memcpy(bb, m2, n * sizeof(float));
    }
    float scale = (float)(1.0 / (48 * spc));               /* rx_path.py:54 */
    amo_moving_average(bb, n, 48 * spc, scale, ma_mode, chunk, avg);
    free(m2);
}

/* ------------------------------------------------------------------ result container */

struct amo_result {
    uint64_t ndet, cap;
    uint64_t* index;
    uint64_t* secs;
    double* frac;
    float* chips;
    amo_frame* frames;
    uint64_t nmsg;
    char** msgs;
    uint64_t calls;
};

static amo_result* res_new(void)
{
    amo_result* r = (amo_result*)calloc(1, sizeof(amo_result));
    return r;
}
static void res_push_det(amo_result* r, uint64_t index, uint64_t secs, double frac, const float* chips)
{
    if (r->ndet == r->cap) {
        r->cap = r->cap ? 2 * r->cap : 256;
        r->index = (uint64_t*)realloc(r->index, r->cap * sizeof(uint64_t));
        r->secs = (uint64_t*)realloc(r->secs, r->cap * sizeof(uint64_t));
        r->frac = (double*)realloc(r->frac, r->cap * sizeof(double));
        r->chips = (float*)realloc(r->chips, r->cap * 240 * sizeof(float));
    }
    r->index[r->ndet] = index;
    r->secs[r->ndet] = secs;
    r->frac[r->ndet] = frac;
    memcpy(r->chips + 240 * r->ndet, chips, 240 * sizeof(float));
    r->ndet++;
}
uint64_t amo_num_det(const amo_result* r) { return r->ndet; }
uint64_t amo_num_calls(const amo_result* r) { return r->calls; }
void amo_get_det(const amo_result* r, uint64_t* index, uint64_t* secs, double* frac, float* chips)
{
    if (index) memcpy(index, r->index, r->ndet * sizeof(uint64_t));
    if (secs) memcpy(secs, r->secs, r->ndet * sizeof(uint64_t));
    if (frac) memcpy(frac, r->frac, r->ndet * sizeof(double));
    if (chips) memcpy(chips, r->chips, r->ndet * 240 * sizeof(float));
}
void amo_get_frames(const amo_result* r, amo_frame* frames)
{
    if (r->frames) memcpy(frames, r->frames, r->ndet * sizeof(amo_frame));
}
uint64_t amo_num_msgs(const amo_result* r) { return r->nmsg; }
const char* amo_msg(const amo_result* r, uint64_t k) { return r->msgs[k]; }
void amo_free(amo_result* r)
{
    if (!r) return;
    for (uint64_t k = 0; k < r->nmsg; k++) free(r->msgs[k]);
    free(r->msgs); free(r->index); free(r->secs); free(r->frac); free(r->chips); free(r->frames);
    free(r);
}

/* ------------------------------------------------------------------ preamble detector */

/* preamble_impl.cc:88-98: double accumulator over chips {0,2,7,9}, samples ascending. */
static double correlate_preamble(const float* in, int samples_per_chip)
{
    static const int preamble_bits[10] = {1, 0, 1, 0, 0, 0, 0, 1, 0, 1};
    double corr = 0.0;
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < samples_per_chip; j++)
            if (preamble_bits[i]) corr += in[i * samples_per_chip + j];
    return corr;
}

/* One general_work() call (preamble_impl.cc:139-246). in/inavg point at the oldest history item.
 * Returns 240 (packet written to out, *i_found set) or 0; *consumed as consume_each(). */
static int general_work_once(const amo_params* p, int mininputs, const float* in, const float* inavg,
                             float* out, int* consumed, int* i_found)
{
    const float spc = p->spc, sps = p->sps, thr = p->threshold;
    /* :147-151 */
    int ninputs = mininputs - (mininputs % (int)spc) - (int)spc;
    if (ninputs < 0) ninputs = 0;
    if (ninputs <= 0) { *consumed = 0; return 0; }

    for (int i = 0; i < ninputs; i++) {                                /* :172 */
        float pulse_threshold = inavg[i] * thr;                        /* :173 */
        if (in[i] > pulse_threshold) {                                 /* :174 */
            if (in[i + 1] > in[i]) continue;                           /* :175 */
            if (in[i + p->po[1]] < pulse_threshold) continue;          /* :177 */
            if (in[i + p->po[2]] < pulse_threshold) continue;          /* :178 */
            if (in[i + p->po[3]] < pulse_threshold) continue;          /* :179 */

            int late, how_late = 0;                                    /* :182-192 */
            do {
                double now_corr = correlate_preamble(in + i, (int)spc);
                double late_corr = correlate_preamble(in + i + 1, (int)spc);
                late = (late_corr > now_corr);
                if (late) { i++; how_late++; }
            } while (late && how_late < spc);

            float avgpeak = (in[i + p->po[0]] + in[i + p->po[1]]      /* :198-201: float adds, /4.0 in double */
                             + in[i + p->po[2]] + in[i + p->po[3]]) / 4.0;
            float space_threshold = inavg[i] + (avgpeak - inavg[i]) / thr; /* :203 */
            int valid_preamble = 1;
            for (int j = 1.5 * sps; j <= 3 * sps; j++)                 /* :205-206 */
                if (in[i + j] > space_threshold) valid_preamble = 0;
            for (int j = 5 * sps; j <= 7.5 * sps; j++)                 /* :207-208 */
                if (in[i + j] > space_threshold) valid_preamble = 0;
            if (!valid_preamble) continue;                             /* :209 */

            if (ninputs - i < 240 * spc) {                             /* :212-216 */
                *consumed = i - 1 > 0 ? i - 1 : 0;
                return 0;
            }
            for (int j = 0; j < 240; j++)                              /* :219-221 */
                out[j] = in[i + (int)(j * spc)] - inavg[i];
            *i_found = i;
            *consumed = i + 240 * spc;                                 /* :237 float -> int */
            return 240;
        }
    }
    *consumed = ninputs;                                               /* :244 */
    return 0;
}

/* tag_to_timestamp (preamble_impl.cc:100-137) with the last rx_time tag at item 0 (or none: stamps 0). */
static uint64_t g_tag_secs = 0;
static double g_tag_frac = 0.0;
void amo_set_start_time(uint64_t secs, double frac) { g_tag_secs = secs; g_tag_frac = frac; }
static void tag_to_timestamp(uint64_t abs_sample_cnt, int rate, uint64_t* secs, double* frac)
{
    uint64_t int_offset = abs_sample_cnt / (uint64_t)rate;                    /* :122 */
    double frac_offset = (abs_sample_cnt % (uint64_t)rate) / (double)rate;    /* :123 */
    uint64_t abs_whole = g_tag_secs + int_offset;                             /* :125 */
    double abs_frac = g_tag_frac + frac_offset;                               /* :126 */
    if (abs_frac > 1.0f) { abs_frac -= 1.0f; abs_whole += 1; }               /* :127-130 */
    *secs = abs_whole;
    *frac = abs_frac;
}

amo_result* amo_scan(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db)
{
    amo_params p;
    amo_make_params(rate, threshold_db, &p);
    amo_result* r = res_new();
    const uint64_t H = (uint64_t)(p.history - 1);
    const uint64_t slack = 64 + (uint64_t)(40.0 * (rate / 2.0e6));
    float* s0 = (float*)calloc(1 + H + n + slack, sizeof(float));
    float* s1 = (float*)calloc(1 + H + n + slack, sizeof(float));
    float* a0 = s0 + 1;
    float* a1 = s1 + 1;
    if (n) {
        memcpy(a0 + H, bb, n * sizeof(float));
        memcpy(a1 + H, avg, n * sizeof(float));
    }
    const uint64_t ntot = n + H;
    uint64_t pos = 0;
    float out[240];
    for (;;) {
        uint64_t remaining = ntot - pos;
        if (remaining > 0x7fffffffull) remaining = 0x7fffffffull;
        int consumed = 0, i_found = 0;
        int ret = general_work_once(&p, (int)remaining, a0 + pos, a1 + pos, out, &consumed, &i_found);
        r->calls++;
        if (ret == 240) {
            uint64_t secs; double frac;
            tag_to_timestamp(pos + (uint64_t)i_found, p.rate_int, &secs, &frac);  /* :224 */
            res_push_det(r, pos + (uint64_t)i_found, secs, frac, out);
        }
        pos += (uint64_t)consumed;
        if (consumed == 0 && ret == 0) break;
    }
    free(s0); free(s1);
    return r;
}

/* ------------------------------------------------------------------ slicer */

/* slicer_impl.cc:67-100 */
static void llslicer(float bit0, float bit1, float ref, int* decision, int* confidence)
{
    float highlimit = ref * 1.414;   /* :71 double product rounded to float */
    float lowlimit = ref * 0.707;    /* :72 */
    int firstchip_inref = ((bit0 > lowlimit) && (bit0 < highlimit));
    int secondchip_inref = ((bit1 > lowlimit) && (bit1 < highlimit));
    if (firstchip_inref && !secondchip_inref) { *decision = 1; *confidence = 1; }
    else if (secondchip_inref && !firstchip_inref) { *decision = 0; *confidence = 1; }
    else if (firstchip_inref && secondchip_inref) { *decision = bit0 > bit1; *confidence = 0; }
    else {
        *decision = bit0 > bit1;
        if (*decision) *confidence = (bit1 < lowlimit * 0.5) ? 1 : 0;   /* :91 compare in double */
        else *confidence = (bit0 < lowlimit * 0.5) ? 1 : 0;             /* :94 */
    }
}

/* slicer_impl.cc:117-182 for one tag. */
void amo_slice_packet(const float* in, amo_frame* f)
{
    memset(f->data, 0, 14);
    memset(f->lowconfbits, 0, 24);
    unsigned numlowconf = 0;
    f->passed = 0;
    f->crc = 0;
    f->df = 0;
    f->ref_level = (in[0] + in[2] + in[7] + in[9]) / 4.0;             /* :128-131 */
    const float ref = f->ref_level;
    const float* d = in + 16;                                          /* :133 */
    unsigned char pkt_hdr = 0;
    for (int j = 0; j < 5; j++) {                                      /* :135-139 */
        int dec, conf;
        llslicer(d[j * 2], d[j * 2 + 1], ref, &dec, &conf);
        if (dec) pkt_hdr += 1 << (4 - j);
    }
    int is_long = (pkt_hdr == 16 || pkt_hdr == 17 || pkt_hdr == 20 || pkt_hdr == 21); /* :140 */
    int packet_length = is_long ? 112 : 56;                            /* :142 */
    f->nbits = (uint8_t)packet_length;
    for (int j = 0; j < packet_length; j++) {                          /* :146-159 */
        int dec, conf;
        llslicer(d[j * 2], d[j * 2 + 1], ref, &dec, &conf);
        if (dec) f->data[j / 8] += 1 << (7 - (j % 8));
        if (!conf) { if (numlowconf < 24) f->lowconfbits[numlowconf++] = (uint8_t)j; }
    }
    f->numlowconf = (uint8_t)numlowconf;
    f->df = (f->data[0] >> 3) & 0x1F;                                  /* :168 (computed early: harmless) */
    int zeroes = 1;                                                    /* :162-166 */
    for (int m = 0; m < 14; m++) if (f->data[m]) zeroes = 0;
    if (zeroes) return;
    if (!is_long && f->df != 11 && numlowconf > 0) return;             /* :170 */
    if (f->df == 11 && numlowconf >= 10) return;                       /* :171 */
    uint32_t crc = amo_crc24(f->data, packet_length / 8 - 3);          /* :173 */
    uint32_t ap = (uint32_t)f->data[packet_length / 8 - 3] << 16       /* :174-176 */
                | (uint32_t)f->data[packet_length / 8 - 2] << 8
                | (uint32_t)f->data[packet_length / 8 - 1];
    crc ^= ap;                                                         /* :177 */
    f->crc = crc;
    if (crc && (f->df == 11 || f->df == 17)) return;                   /* :182 */
    f->passed = 1;
}

/* slicer_impl.cc:186-192 */
int amo_format_message(const amo_frame* f, int first, char* buf, size_t buflen)
{
    int o = 0;
    for (int m = 0; m < f->nbits / 8; m++) o += snprintf(buf + o, buflen - o, "%02x", (unsigned)f->data[m]);
    /* operator<<(float) prints through %g with the stream precision: 6 until :192 has run once */
    o += snprintf(buf + o, buflen - o, " %06lx %.*g %llu %.10g", (unsigned long)f->crc, first ? 6 : 10,
                  (double)f->ref_level, (unsigned long long)f->secs, f->frac);
    return o;
}

static void slice_all(amo_result* r)
{
    r->frames = (amo_frame*)calloc(r->ndet ? r->ndet : 1, sizeof(amo_frame));
    r->msgs = (char**)calloc(r->ndet ? r->ndet : 1, sizeof(char*));
    r->nmsg = 0;
    for (uint64_t k = 0; k < r->ndet; k++) {
        amo_frame* f = &r->frames[k];
        amo_slice_packet(r->chips + 240 * k, f);
        f->index = r->index[k];
        f->secs = r->secs[k];
        f->frac = r->frac[k];
        if (f->passed) {
            char buf[160];
            amo_format_message(f, r->nmsg == 0, buf, sizeof buf);
            r->msgs[r->nmsg++] = strdup(buf);
        }
    }
}

amo_result* amo_run_streams(const float* bb, const float* avg, uint64_t n, float rate, float threshold_db)
{
    amo_result* r = amo_scan(bb, avg, n, rate, threshold_db);
    slice_all(r);
    return r;
}

amo_result* amo_run_iq(const float* iq, uint64_t n, float rate, float threshold_db, int use_pmf,
                       int ma_mode, int chunk)
{
    float* bb = (float*)malloc((n ? n : 1) * sizeof(float));
    float* avg = (float*)malloc((n ? n : 1) * sizeof(float));
    amo_frontend(iq, n, rate, use_pmf, ma_mode, chunk, bb, avg);
    amo_result* r = amo_run_streams(bb, avg, n, rate, threshold_db);
    free(bb); free(avg);
    return r;
}

/* the whole rx_path incl. the optional DC blocker in front of the demodulator (rx_path.py:39-41) */
amo_result* amo_run_iq_dc(const float* iq, uint64_t n, float rate, float threshold_db, int use_pmf, int use_dcblock,
                          int ma_mode, int chunk)
{
    if (!use_dcblock) return amo_run_iq(iq, n, rate, threshold_db, use_pmf, ma_mode, chunk);
    const int spc = (int)((double)rate / 2e6);
    float* y = (float*)malloc(2 * (n ? n : 1) * sizeof(float));
    amo_dc_blocker(iq, n, 100 * spc, ma_mode == AMO_MA_GR_FLOAT ? AMO_MA_GR_FLOAT : AMO_MA_CANONICAL, y);
    amo_result* r = amo_run_iq(y, n, rate, threshold_db, use_pmf, ma_mode, chunk);
    free(y);
    return r;
}

amo_result* amo_run_slicer(const float* chips, uint64_t ndet, const uint64_t* secs, const double* frac)
{
    amo_result* r = res_new();
    for (uint64_t k = 0; k < ndet; k++) res_push_det(r, 0, secs[k], frac[k], chips + 240 * k);
    slice_all(r);
    return r;
}
