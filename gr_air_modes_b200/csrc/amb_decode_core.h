// Per-message decode arithmetic of SURVEY.md 8 row f4: what the reference's python/parse.py, python/altitude.py and
// python/cpr.py compute for one queued message, as __host__ __device__ functions. The CUDA kernels in amb_decode.cu
// are the only product code that calls them; tests/decode_host_shim.cc compiles the same functions for the host so
// that the arithmetic is checked against the reference's golden where no GPU is present (a test harness, not a
// product path). Citations are file:line under gr-air-modes/python.
#pragma once
#include <math.h>
#include <stdint.h>
#include "../../include/airmodes_b200.h"

#if defined(__CUDACC__)
#define AMB_HD __host__ __device__ __forceinline__
#else
#define AMB_HD static inline
#endif

#define AMB_NO_KEY 0xFFFFFFFFu
#define AMB_NL_MAX 64            // nl_T[k]: smallest |lat| at which cpr.py:48-51 evaluates below k (k = 3..nl_T[0])

// A CPR report as the pairing stage sees it (cpr.py:214-221: [encoded_lat, encoded_lon, time] per icao, per
// even/odd, per airborne/surface list).
struct AmbPosRec {
    uint32_t key;        // (icao24 << 1) | surface; AMB_NO_KEY = not a position message
    uint32_t lat, lon;   // 17-bit encoded values
    uint32_t fmt;        // 0 even, 1 odd ("cpr" field, parse.py:126-127)
    double t;            // secs + frac of the message
};

// What the pairing stage found for a position message: the latest even and odd report of its aircraft that are
// still alive (cpr.py:196-204), and which is newer (cpr.py:229).
struct AmbPair {
    uint32_t elat, elon, olat, olon;
    int have, mostrecent;
};

// ---- bit fields: data_field.get_bits (parse.py:71-87), fields 1-based from the MSB ---------------------------------
struct AmbMsg { uint64_t w0, w1; int numbits; };   // message bits 1..64 in w0 (MSB first), 65..112 in the top of w1

AMB_HD uint64_t amb_bits(const AmbMsg& m, int s, int n)
{
    if (s + n - 1 > m.numbits) return 0;            // negative shift -> ValueError -> field reads 0 (parse.py:83-86)
    const int a = s - 1;
    uint64_t x;
    if (a == 0) x = m.w0;
    else if (a < 64) x = (m.w0 << a) | (m.w1 >> (64 - a));
    else x = m.w1 << (a - 64);
    return x >> (64 - n);
}

AMB_HD AmbMsg amb_msg_from_frame(const amb_frame& f)
{
    AmbMsg m;
    uint64_t hi = 0, lo = 0;
    for (int k = 0; k < 7; k++) hi = (hi << 8) | f.data[k];
    if (f.nbits == 112) {
        for (int k = 7; k < 14; k++) lo = (lo << 8) | f.data[k];
        // modes_reply.is_long: int(data, 16) > (1 << 56) (parse.py:222-223)
        if (hi > 1 || (hi == 1 && lo > 0)) {
            m.w0 = (hi << 8) | (lo >> 48); m.w1 = lo << 16; m.numbits = 112;
        } else {                                    // a 28-digit string with a small value parses as a short reply
            m.w0 = lo << 8; m.w1 = 0; m.numbits = 56;
        }
    } else {
        m.w0 = hi << 8; m.w1 = 0; m.numbits = 56;
    }
    return m;
}

// ---- altitude.py --------------------------------------------------------------------------------------------------
AMB_HD int amb_gray2bin(int g)                      // altitude.py:110-117
{
    for (int i = g >> 1; i != 0; i >>= 1) g ^= i;
    return g;
}

// decode_alt(alt, bit13) (altitude.py:28-108). Returns 0 and *feet, or 1 for MetricAltError (:32-43).
AMB_HD int amb_decode_alt(int alt, int bit13, int32_t* feet)
{
    if ((alt & 0x40) && bit13) return 1;
    if (alt & 0x10) {                               // Mode S style, 25 ft (:45-56)
        const int t = bit13 ? (((alt & 0x3F80) >> 2) | ((alt & 0x20) >> 1)) : ((alt & 0x1FE0) >> 1);
        *feet = ((alt & 0x0F) | t) * 25 - 1000;
        return 0;
    }
    if (!bit13) alt = (alt & 0x3F) | (alt & (0x0FC0 << 1));      // :67 exactly as written (operator precedence)
    const int big = ((alt & 0x0002) >> 1) + ((alt & 0x0008) >> 2) + ((alt & 0x0020) >> 3) + ((alt & 0x0080) >> 4) +
                    ((alt & 0x0200) >> 5) + ((alt & 0x0800) >> 6) + ((alt & 0x0001) << 6) + ((alt & 0x0004) << 5);   // :82-89
    int d = amb_gray2bin(big);
    int c = amb_gray2bin(((alt & 0x0100) >> 8) + ((alt & 0x0400) >> 9) + ((alt & 0x1000) >> 10));               // :95-96
    if (c == 7) c = 5;
    if (d % 2) c = 6 - c;
    *feet = d * 500 + c * 100 - 1300;               // :104-106
    return 0;
}

AMB_HD int amb_decode_id(int v)                     // parse.py:233-254
{
    const int a = ((v & 0x0800) >> 11) + ((v & 0x0200) >> 8) + ((v & 0x0080) >> 5);
    const int b = ((v & 0x0020) >> 5) + ((v & 0x0008) >> 2) + ((v & 0x0002) << 1);
    const int c = ((v & 0x1000) >> 12) + ((v & 0x0400) >> 9) + ((v & 0x0100) >> 6);
    const int d = ((v & 0x0010) >> 2) + ((v & 0x0004) >> 1) + ((v & 0x0001) << 2);
    return a * 1000 + b * 100 + c * 10 + d;
}

AMB_HD void amb_ident48(uint64_t v, char* out)      // charmap + parseBDS08 / parseMB_id (parse.py:257-279, 374-378)
{
    for (int i = 0; i < 8; i++) {
        const int d = (int)((v >> (42 - 6 * i)) & 0x3F);
        char ch = ' ';
        if (d > 0 && d < 27) ch = (char)('A' + d - 1);
        else if (d > 47 && d < 58) ch = (char)('0' + d - 48);
        out[i] = ch;
    }
}

// ---- one message -> fields (everything but the CPR resolution) ------------------------------------------------------
AMB_HD void amb_fields_blank(amb_fields* r, int df, uint32_t ecc)
{
    r->icao = ecc; r->ecc = ecc;
    r->df = (uint8_t)df; r->status = 0; r->bds = 0; r->subtype = 0xFF;
    r->ca = r->fs = r->vs = r->ri = 0; r->sl = r->cc = r->dr = r->um = 0;
    r->ftc = r->cat = r->cpr_format = r->surface = 0; r->eps = r->ast = r->bds2 = r->tti = 0;
    r->altitude = AMB_NO_ALTITUDE; r->squawk = -1; r->threat_alt = AMB_NO_ALTITUDE;
    r->cpr_lat = r->cpr_lon = 0;
    for (int k = 0; k < 4; k++) { r->aux[k] = 0; r->val[k] = NAN; }
    for (int k = 0; k < 8; k++) r->ident[k] = 0;
    r->lat = r->lon = r->range = r->bearing = NAN;
    r->pad_ = 0;
}

AMB_HD void amb_alt_into(amb_fields* r, int code, int bit13, int32_t* dst, int flag = AMB_FS_METRIC_ALT)
{
    int32_t feet;
    if (amb_decode_alt(code, bit13, &feet)) r->status |= (uint8_t)flag;
    else *dst = feet;
}

// The reference's parser raises NoHandlerError while building the field dict, so nothing of the message survives
// (parse.py:52-68, make_parser :431-434): keep only what identifies it.
AMB_HD void amb_no_handler(amb_fields* r, int keep_icao)
{
    const uint32_t icao = r->icao, ecc = r->ecc; const int df = r->df;
    amb_fields_blank(r, df, ecc);
    if (keep_icao) r->icao = icao;
    r->status = AMB_FS_NO_HANDLER;
}

// modes_reply / me_reply / bds09_reply / mb_reply / tcas_reply field tables (parse.py:89-220) and the parseBDS*
// functions (parse.py:276-372). pos->key stays AMB_NO_KEY unless the message is a BDS0,5 / BDS0,6 position.
AMB_HD void amb_decode_fields(const amb_frame& f, amb_fields* r, AmbPosRec* pos)
{
    pos->key = AMB_NO_KEY; pos->lat = pos->lon = 0; pos->fmt = 0;
    pos->t = (double)f.secs + f.frac;
    const AmbMsg g = amb_msg_from_frame(f);
    const int df = (int)amb_bits(g, 1, 5);
    amb_fields_blank(r, df, f.crc);
    if (!f.passed) { r->status = AMB_FS_NOT_QUEUED; return; }
    if (!(df == 0 || df == 4 || df == 5 || df == 11 || df == 16 || df == 17 || df == 20 || df == 21 || df == 24)) {
        r->status = AMB_FS_NO_HANDLER;              // parse.py:210-220 has no table for this DF
        return;
    }
    if (df == 0 || df == 16) {
        r->vs = (uint8_t)amb_bits(g, 6, 1); r->sl = (uint8_t)amb_bits(g, 9, 3); r->ri = (uint8_t)amb_bits(g, 14, 4);
        if (df == 0) r->cc = (uint8_t)amb_bits(g, 7, 1);
        amb_alt_into(r, (int)amb_bits(g, 20, 13), 1, &r->altitude);          // msprint.py:64,251
    } else if (df == 4 || df == 5 || df == 20 || df == 21) {
        r->fs = (uint8_t)amb_bits(g, 6, 3); r->dr = (uint8_t)amb_bits(g, 9, 5); r->um = (uint8_t)amb_bits(g, 14, 6);
        if (df >= 20) {
            // "mb": (33,56, mb_reply) (parse.py:218-219): the sub-fields are read from that 56-bit word, which is 0 when
            // the reply is short (negative shift, parse.py:83-86). b(s, n) = mb_reply/tcas_reply field at message bit s.
            AmbMsg mb; mb.w0 = amb_bits(g, 33, 56) << 8; mb.w1 = 0; mb.numbits = 56;
#define AMB_MB(s, n) amb_bits(mb, (s) - 32, (n))
            const int bds1 = (int)AMB_MB(33, 4), bds2 = (int)AMB_MB(37, 4);
            r->bds = (uint8_t)bds1; r->bds2 = (uint8_t)bds2;
            if (bds1 > 3 || bds2 != 0) { amb_no_handler(r, 0); return; }     // parse.py:185-190
            if (bds1 == 1) {
                r->aux[0] = (uint32_t)AMB_MB(45, 20); r->aux[1] = (uint32_t)AMB_MB(65, 16);
                r->aux[2] = (uint32_t)AMB_MB(81, 8);  r->aux[3] = (uint32_t)AMB_MB(41, 4);   // acs, bcs, ecs, cfs
            } else if (bds1 == 2) {
                amb_ident48(AMB_MB(41, 48), r->ident);                      // parse.py:374-378
            } else if (bds1 == 3) {
                const int tti = (int)AMB_MB(61, 2);
                r->tti = (uint8_t)tti;
                if (tti == 3) { amb_no_handler(r, 0); return; }              // tcas_reply has types 0-2 (parse.py:157-165)
                r->aux[0] = (uint32_t)AMB_MB(41, 14); r->aux[1] = (uint32_t)AMB_MB(55, 4);
                r->aux[2] = (uint32_t)(AMB_MB(59, 1) | (AMB_MB(60, 1) << 1));   // rat | mte << 1
                if (tti == 1) r->aux[3] = (uint32_t)AMB_MB(63, 26);          // tid
                else if (tti == 2) {
                    r->aux[3] = (uint32_t)(AMB_MB(76, 7) | (AMB_MB(83, 6) << 8));   // tidr | tidb << 8
                    amb_alt_into(r, (int)AMB_MB(63, 13), 1, &r->threat_alt, AMB_FS_METRIC_THREAT);   // parse.py:407
                }
            }
#undef AMB_MB
        }
        if (df == 4 || df == 20) amb_alt_into(r, (int)amb_bits(g, 20, 13), 1, &r->altitude);
        else r->squawk = amb_decode_id((int)amb_bits(g, 20, 13));
    } else if (df == 11) {
        r->ca = (uint8_t)amb_bits(g, 6, 3); r->icao = (uint32_t)amb_bits(g, 9, 24);
    } else if (df == 17) {
        r->ca = (uint8_t)amb_bits(g, 6, 3); r->icao = (uint32_t)amb_bits(g, 9, 24);
        AmbMsg m; m.w0 = amb_bits(g, 33, 56) << 8; m.w1 = 0; m.numbits = 56;       // "me" (parse.py:217)
        const int ftc = (int)amb_bits(m, 1, 5);
        r->ftc = (uint8_t)ftc;
        if (ftc >= 1 && ftc <= 4) {                                         // parse.py:140-152 maps ftc to the BDS register
            r->bds = 0x08;
            r->cat = (uint8_t)amb_bits(m, 6, 3);
            amb_ident48(amb_bits(m, 9, 48), r->ident);
        } else if (ftc >= 5 && ftc <= 8) {
            r->bds = 0x06; r->surface = 1;
            r->cpr_format = (uint8_t)amb_bits(m, 22, 1);
            r->cpr_lat = (uint32_t)amb_bits(m, 23, 17); r->cpr_lon = (uint32_t)amb_bits(m, 40, 17);
            r->val[0] = (double)(int)amb_bits(m, 14, 7) * 360. / 128;        // parse.py:283
            pos->key = (r->icao << 1) | 1u;
        } else if (ftc >= 9 && ftc <= 18 && ftc != 15) {
            r->bds = 0x05;
            r->cpr_format = (uint8_t)amb_bits(m, 22, 1);
            r->cpr_lat = (uint32_t)amb_bits(m, 23, 17); r->cpr_lon = (uint32_t)amb_bits(m, 40, 17);
            amb_alt_into(r, (int)amb_bits(m, 9, 12), 0, &r->altitude);       // parse.py:277
            pos->key = (r->icao << 1);
        } else if (ftc == 19) {
            r->bds = 0x09;
            const int sub = (int)amb_bits(m, 6, 3);
            if (sub == 0) {                                                 // parseBDS09_0 (parse.py:288-313)
                r->subtype = 0;
                int vs = (int)amb_bits(m, 42, 9) * 32;
                if (amb_bits(m, 41, 1)) vs = 0 - vs;
                double tr = (double)((int)amb_bits(m, 35, 6) * 15) / 62;
                if (amb_bits(m, 34, 1)) tr = 0 - tr;
                int ns = (int)amb_bits(m, 23, 11) - 1, ew = (int)amb_bits(m, 11, 11) - 1;
                const double vel = hypot((double)ns, (double)ew);
                if (amb_bits(m, 10, 1)) ew = 0 - ew;
                if (amb_bits(m, 22, 1)) ns = 0 - ns;
                double hdg = atan2((double)ew, (double)ns) * (180.0 / M_PI);
                if (hdg < 0) hdg += 360;
                r->val[0] = vel; r->val[1] = hdg; r->val[2] = (double)vs; r->val[3] = tr;
            } else if (sub == 1 || sub == 2) {                              // parseBDS09_1 (parse.py:315-348)
                r->subtype = 1;
                int geo = (int)amb_bits(m, 50, 6) * 25;
                if (amb_bits(m, 49, 1)) geo = 0 - geo;
                double vs = (double)((int)amb_bits(m, 38, 9) - 1) * 64;
                if (amb_bits(m, 37, 1)) vs = 0 - vs;
                double ns = (double)(int)amb_bits(m, 26, 10), ew = (double)(int)amb_bits(m, 15, 10);
                if (sub == 2) { ns *= 4; ew *= 4; }
                const double vel = hypot(ns, ew);
                if (amb_bits(m, 14, 1)) ew = 0 - ew;
                double hdg = (ns == 0) ? 0.0 : atan(ew / ns) * (180.0 / M_PI);
                if (amb_bits(m, 25, 1)) hdg = 180 - hdg;
                if (hdg < 0) hdg += 360;
                r->val[0] = vel; r->val[1] = hdg; r->val[2] = vs; r->val[3] = (double)geo;
            } else if (sub == 3 || sub == 4) {                              // parseBDS09_3 (parse.py:350-364)
                r->subtype = 3;
                r->ast = (uint8_t)amb_bits(m, 25, 1);
                int vel = (int)amb_bits(m, 26, 10);
                if (sub == 4) vel *= 4;
                double vs = (double)((int)amb_bits(m, 38, 9) - 1) * 64;
                if (amb_bits(m, 37, 1) == 1) vs = 0 - vs;
                r->val[0] = (double)(int)amb_bits(m, 14, 1) * 360. / 1024;   // :353 reads "mhs", the 1-bit status field
                r->val[1] = (double)vel; r->val[2] = vs;
                r->val[3] = (double)((int)amb_bits(m, 50, 6) - 1) * 25;
            } else {                                                        // bds09_reply.get_type() -> None (parse.py:110-117)
                amb_no_handler(r, 1); return;
            }
        } else if (ftc == 28) {
            r->bds = 0x61;
            r->eps = (uint8_t)amb_bits(m, 9, 3);
        } else {
            amb_no_handler(r, 1); return;
        }
        if (pos->key != AMB_NO_KEY) { pos->lat = r->cpr_lat; pos->lon = r->cpr_lon; pos->fmt = r->cpr_format; }
    }
    /* df == 24: "ke"/"nd"/"md" have no consumer in the reference */
}

// ---- cpr.py ---------------------------------------------------------------------------------------------------------
// nl(declat_in) (cpr.py:46-51) through the transition table built by amb_build_nl_table() from the same libm
// expression: the value is floor(f(|lat|)) with f decreasing, so NL = the largest k with |lat| < T[k].
AMB_HD int amb_nl(double lat, const double* T)
{
    const double a = fabs(lat);
    if (a >= 87.0) return 1;
    for (int k = (int)T[0]; k >= 3; k--)
        if (a < T[k]) return k;
    return 2;
}

AMB_HD long long amb_pymod(long long a, long long b)     // Python's % for b > 0
{
    const long long r = a % b;
    return r < 0 ? r + b : r;
}

// cpr_resolve_global (cpr.py:89-153). 0 = ok, AMB_FS_CPR_NO_POS, or AMB_FS_CPR_NO_POS | AMB_FS_CPR_STRADDLE.
AMB_HD int amb_cpr_global(const AmbPair& pr, int surface, int have_loc, double mylat, double mylon, const double* T,
                          double* out_lat, double* out_lon)
{
    if (surface && !have_loc) return AMB_FS_CPR_NO_POS;                     // :97-99
    const double span = surface ? 90.0 : 360.0;                             // dlat/dlon (:37-45, :53-59)
    const double dle = span / 60, dlo = span / 59;
    const double elat = (double)pr.elat, elon = (double)pr.elon, olat = (double)pr.olat, olon = (double)pr.olon;
    const long long j = (long long)floor(((59 * elat - 60 * olat) / 131072) + 0.5);                 // :107
    double rle = dle * ((double)amb_pymod(j, 60) + elat / 131072);          // :109
    double rlo = dlo * ((double)amb_pymod(j, 59) + olat / 131072);          // :110
    if (rle > 270.0) rle -= 360.0;
    if (rlo > 270.0) rlo -= 360.0;
    if (amb_nl(rle, T) != amb_nl(rlo, T)) return AMB_FS_CPR_NO_POS | AMB_FS_CPR_STRADDLE;           // :120-121
    double rlat = pr.mostrecent ? rlo : rle;
    if (surface && mylat < 0) rlat -= 90;                                   // :129-131
    const int n = amb_nl(rlat, T);
    const int nn = (n - pr.mostrecent) > 1 ? (n - pr.mostrecent) : 1;
    const double dl = span / nn;                                            // dlon(rlat, mostrecent, surface) :133
    const long long m = (long long)floor(((elon * (n - 1) - olon * n) / 131072) + 0.5);             // :136
    const double enclon = pr.mostrecent ? olon : elon;
    double rlon = dl * ((double)amb_pymod(m, nn) + enclon / 131072.);       // :147
    if (surface) {                                                          // :152-158, `zone` with Python-3 true division
        double wat = mylon;
        if (wat < 0) wat += 360;
        const double za = 90 * ((double)(long long)wat / 90), zb = 90 * ((double)(long long)rlon / 90);
        rlon += (za - zb);
    }
    if (rlon > 180) rlon -= 360.0;                                          // :160-162
    *out_lat = rlat; *out_lon = rlon;
    return 0;
}

AMB_HD void amb_range_bearing(double a_lat, double a_lon, double b_lat, double b_lon, double* rnge, double* bearing)
{                                                                           // cpr.py:158-181
    const double esquared = (1 / 298.257223563) * (2 - (1 / 298.257223563));
    const double earth_radius_mi = 3963.19059 * (M_PI / 180);
    const double delta_lat = b_lat - a_lat, delta_lon = b_lon - a_lon;
    const double avg_lat = ((a_lat + b_lat) / 2.0) * M_PI / 180;
    const double s2 = pow(sin(avg_lat), 2.0);
    const double R1 = earth_radius_mi * (1.0 - esquared) / pow((1.0 - esquared * s2), 1.5);
    const double R2 = earth_radius_mi / sqrt(1.0 - esquared * s2);
    const double north = R1 * delta_lat;
    const double east = R2 * cos(avg_lat) * delta_lon;
    double b = atan2(east, north) * (180.0 / M_PI);
    if (b < 0.0) b += 360.0;
    *rnge = hypot(east, north);
    *bearing = b;
}

// parseBDS05 / parseBDS06 tail (parse.py:276-286) + cpr_decoder.decode after the pair lookup (cpr.py:226-240).
AMB_HD void amb_resolve_position(amb_fields* r, const AmbPair& pr, int have_loc, double mylat, double mylon, const double* T)
{
    if (!pr.have) { r->status |= AMB_FS_CPR_NO_POS; return; }               // cpr.py:231
    double lat, lon;
    const int e = amb_cpr_global(pr, r->surface, have_loc, mylat, mylon, T, &lat, &lon);
    if (e) { r->status |= (uint8_t)e; return; }
    r->lat = lat; r->lon = lon; r->status |= AMB_FS_HAS_POS;
    if (have_loc) {
        amb_range_bearing(mylat, mylon, lat, lon, &r->range, &r->bearing);
        r->status |= AMB_FS_HAS_RANGE;
    }
}

// Latest-report bookkeeping of cpr_decoder.decode for ONE message given the table state before it (cpr.py:214-229):
// used by the pairing kernel lane by lane and, sequentially, by the host shim.
// A stored report stamped LATER than the current message can only stem from an earlier stream (frame timestamps restart
// after amb_reset / a new recording, while a decoder keeps its table): the reference's wall clock would have expired it
// long ago, so it counts as expired here too instead of being paired with.
AMB_HD int amb_report_alive(double now, double t, int surface) { const double age = now - t; return age >= 0.0 && !(age > (surface ? 25.0 : 10.0)); }

// `me` has just been stored as the latest report of its format; (o_*) is the latest stored report of the other
// format, if any. -> what cpr_decoder.decode hands to cpr_resolve_global (cpr.py:226-229), or have = 0 (:231).
AMB_HD AmbPair amb_make_pair(const AmbPosRec& me, int o_have, uint32_t o_lat, uint32_t o_lon, double o_t)
{
    AmbPair pr; pr.have = 0; pr.mostrecent = 0; pr.elat = pr.elon = pr.olat = pr.olon = 0;
    if (o_have && amb_report_alive(me.t, o_t, (int)(me.key & 1u))) {       // weed_poslists (cpr.py:196-204)
        pr.have = 1;
        if (me.fmt) { pr.olat = me.lat; pr.olon = me.lon; pr.elat = o_lat; pr.elon = o_lon; pr.mostrecent = (me.t - o_t) > 0 ? 1 : 0; }
        else        { pr.elat = me.lat; pr.elon = me.lon; pr.olat = o_lat; pr.olon = o_lon; pr.mostrecent = (o_t - me.t) > 0 ? 1 : 0; }
    }
    return pr;
}

// Transition latitudes of nl() from the host libm, evaluated exactly as cpr.py:48-51 writes it (float ** 2 and
// float ** -1 are C pow() in CPython). T[0] = nl(0) = the largest NL; T[k], k = 3..T[0]. Host only: called once by
// amb_decoder_create, the table then lives in device memory.
static inline double amb_nl_formula(double x)
{
    const double c = cos(M_PI / (2.0 * 15));
    const double a = acos(1.0 - (1.0 - c) / pow(cos((M_PI / 180.0) * fabs(x)), 2.0));
    return floor((2.0 * M_PI) * pow(a, -1.0));
}
static inline void amb_build_nl_table(double* T)
{
    for (int k = 0; k < AMB_NL_MAX; k++) T[k] = 0.0;
    const int kmax = (int)amb_nl_formula(0.0);
    T[0] = (double)(kmax < AMB_NL_MAX ? kmax : AMB_NL_MAX - 1);
    for (int k = 3; k <= (int)T[0]; k++) {
        double lo = 0.0, hi = 86.99;                 // f(lo) >= k, f(hi) = 2 < k
        for (;;) {
            const double mid = lo + (hi - lo) / 2;
            if (!(mid > lo && mid < hi)) break;
            if (amb_nl_formula(mid) >= (double)k) lo = mid; else hi = mid;
        }
        T[k] = hi;
    }
}
