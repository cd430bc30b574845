// Host-side derivation of AmbParams from (rate, threshold): what preamble_impl::set_rate / set_threshold and the loop
// headers of general_work compute. Used by amb_api.cu; a header so that tests/simt can drive the kernels with the
// very same numbers.
#pragma once
#include <math.h>
#include <string.h>
#include "amb_internal.h"

// ---- parameters: preamble_impl.cc:56-68,158-162,192,205-208,212,237; rx_path.py:35,49,54 -------------------
static inline int compute_params(float channel_rate, float threshold_db, int use_pmf, AmbParams* P, int* chip_off)
{
    const int chip_rate = 2000000;                       // preamble_impl.cc:46
    memset(P, 0, sizeof *P);
    P->spc_f = channel_rate / chip_rate;                 // :57
    P->sps_f = P->spc_f * 2;                             // :58
    if (!(P->spc_f >= 1.0f) || P->spc_f > (float)AMB_MAX_SPC) return AMB_ERR_RATE;
    P->spc_i = (int)P->spc_f;
    const int spc_rx = (int)((double)channel_rate / 2e6);   // rx_path.py:35
    if (spc_rx != P->spc_i) return AMB_ERR_RATE;
    P->rate_int = (int)channel_rate;                     // :60
    P->L = 48 * spc_rx;                                  // rx_path.py:54
    P->H = (int)(unsigned)P->sps_f - 1;                  // :62 set_history
    P->po1 = (int)(2 * P->spc_f); P->po2 = (int)(7 * P->spc_f); P->po3 = (int)(9 * P->spc_f);   // :158-162
    {   // loop bounds exactly as the for statements evaluate them (:205, :207)
        int j = 1.5 * P->sps_f; P->qa0 = j; P->qa1 = j - 1;
        for (; j <= 3 * P->sps_f; j++) P->qa1 = j;
        j = 5 * P->sps_f; P->qb0 = j; P->qb1 = j - 1;
        for (; j <= 7.5 * P->sps_f; j++) P->qb1 = j;
    }
    {   // do { ...; if (late) how_late++; } while (late && how_late < spc)  (:184-192) with late always true
        int how_late = 0;
        do { how_late++; } while (how_late < P->spc_f);
        P->maxlate = how_late;
    }
    P->skip_f = 240 * P->spc_f;                          // :212, :237
    P->skip0 = (int)P->skip_f;
    P->thr = powf(10., threshold_db / 20.);              // :67
    P->scale_p = (float)(1.0 / spc_rx);                  // rx_path.py:49
    P->scale_a = (float)(1.0 / (48 * spc_rx));           // rx_path.py:54
    P->use_pmf = use_pmf ? 1 : 0;
    for (int j = 0; j < 240; j++) chip_off[j] = (int)(j * P->spc_f);   // :220
    const double eps = 1.0 / 32768.0;
    float cT = (float)((double)P->thr * (double)P->scale_a * (1.0 - eps) * (1.0 - eps));
    P->cT = nextafterf(cT, 0.0f);
    P->one_eps = 1.0f + (float)eps;
    P->gfac = 1.0f / 524288.0f;                          // 2^-19
    int fwd = 10 * P->spc_i;
    if (P->qb1 > fwd) fwd = P->qb1;
    if (P->po3 > fwd) fwd = P->po3;
    P->fwd = fwd + 2;
    if (P->skip_f == floorf(P->skip_f)) {
        P->i_exact = (1ll << 24) + 1 - P->skip0;         // first i with i+skip odd above 2^24 rounds away
    } else {                                             // fractional 240*spc: find the first i where :237 rounds up
        long long i = 0;
        for (; i < (1ll << 25); i++) if ((int)((float)i + P->skip_f) != (int)i + P->skip0) break;
        P->i_exact = i;
    }
    if (P->maxlate > 14) return AMB_ERR_RATE;
    return AMB_OK;
}
