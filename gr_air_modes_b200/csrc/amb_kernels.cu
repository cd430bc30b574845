// sm_100a kernels of the Mode S receive hot path. Citations are file:line under gr-air-modes.
//
// Pipeline per amb_process() call (all on one CUDA stream, nothing round-trips to the host):
//   scan     amb_scan_kernel      streaming pass over IQ (8 B/sample, the HBM-bound kernel): |x|^2, pulse
//                                 matched filter, noise-floor window and the first four pulse tests of
//                                 preamble_impl.cc:173-179 as a CONSERVATIVE fp32 pre-filter -> candidate bitmap
//   compact  amb_compact_kernel   bitmap -> ordered candidate list
//   exact    amb_exact_kernel     one warp per candidate: canonical arithmetic (bit-for-bit the oracle's),
//                                 re-checks :173-179, late-gate :182-192, quiet zones :198-209
//   resolve  amb_walk_*           reproduces the visit order of the sequential scan loop (:172, :190, :209,
//                                 :212-216, :237) over the sparse candidate list
//   slice    amb_slice_kernel     one warp per accepted preamble: 240-chip extraction :219-221, llslicer
//                                 slicer_impl.cc:67-100, packet rules :117-182, CRC modes_crc.cc:55-63
#include "amb_internal.h"

#define FULL 0xffffffffu

__constant__ int c_chip_off[240];        // int(j*spc) (preamble_impl.cc:220)
__constant__ unsigned int c_crc_rem[96]; // x^(t+24) mod 0xFFF409, t = distance of a message bit from the parity field

// ------------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + TMA 1-D bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ const float2* seg_ptr(const AmbSegs& S, int j) {
    if (j < S.n_carry) return S.carry + j;
    j -= S.n_carry;
    if (j < S.n_main) return S.main_ + j;
    return S.tail + (j - S.n_main);
}

// ------------------------------------------------------------------------------------------------
// scan kernel
// ------------------------------------------------------------------------------------------------
// Work decomposition: the evaluated range is cut into contiguous spans of rows_per_span rows (a row =
// 128 samples, 4 per lane); ONE WARP owns a span and streams through it on its own, with a private
// 4-stage TMA ring (lane 0 issues cp.async.bulk of 2 KiB stages, all lanes wait on the stage mbarrier).
// No block-level barrier exists in this kernel; a CTA is just four independent warps.
//
// Arithmetic (deliberately NOT the canonical one - this is a filter, the exact stage decides):
//   m2 = re*re + im*im (fma), bbs = sum of the last SPC m2 (all-positive adds, unscaled PMF),
//   Pr = row-local inclusive prefix of bbs (warp scan), Rt = row total,
//   W[n] = sum of the last L bbs = (A - Pr_kd[posd]) + Pr_k[pos] with A = totals of the rows in between.
//   Exact test (preamble_impl.cc:173-174) is  bb > fl(avg*T)  with bb = fl(S1)*sp, avg = fl(sum bb)*sa;
//   since sp cancels, it is implied by  bbs >= cT*W - G  with cT = T*sa*(1-eps)^2 and the absolute guard
//   G = cT*gfac*(Rt+A) >= cT * (accumulated rounding of W)  (every operand of W is <= Rt+A, <= 24 roundings
//   of 2^-24 each; gfac = 2^-19).  The other tests use the same lowered threshold; the peak test
//   in[i+1] > in[i] (:175) is relaxed by (1+eps). eps = 2^-15 dwarfs the <= 2^-20 relative error of bbs.
//   Result: a superset of the reference's candidates; typically < 0.01 % extra.
template <int SPC, bool PMF> struct ScanCfg {
    static constexpr int FL = PMF ? SPC : 1;       // pulse-matched-filter length (rx_path.py:48-51)
    static constexpr int L = 48 * SPC;             // noise-floor window (rx_path.py:54)
    static constexpr int RB = L / AMB_ROW;
    static constexpr int LMOD = L % AMB_ROW;
    static constexpr int PRR = (RB + 2 <= 2) ? 2 : ((RB + 2 <= 4) ? 4 : 8);
    static constexpr int NST = 4;
    static constexpr int WARM = (RB + 2 + 1) & ~1;
    static constexpr int WARP_BYTES = NST * 2048 + 1024 + 1024 + PRR * 512 + 64;
};

__device__ __forceinline__ uint32_t spread8(uint32_t x) {  // 8 bits -> bit positions 0,4,...,28
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

template <int SPC, bool PMF>
__global__ void __launch_bounds__(128) amb_scan_kernel(const __grid_constant__ AmbScanArgs a)
{
    using C = ScanCfg<SPC, PMF>;
    constexpr int FL = C::FL;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span = blockIdx.x * 4 + warp;
    if (span >= a.n_spans) return;

    unsigned char* ws = smem + warp * C::WARP_BYTES;
    float* m2r = reinterpret_cast<float*>(ws + C::NST * 2048);
    float* bbr = m2r + 256;
    float* prr = bbr + 256;
    const uint32_t iq_s = smem_u32(ws);
    const uint32_t bar0 = smem_u32(prr + C::PRR * 128);

    const int ra = a.row_lo + span * a.rows_per_span;      // evaluate rows [ra, rb)
    const int rb = min(ra + a.rows_per_span, a.row_hi);
    const int rs = max(ra - C::WARM, 0);                   // first row computed (warm-up of the windows)
    const int g0 = rs >> 1;
    const int nstages = (rb >> 1) - g0 + 1;                // row rb is computed as look-ahead only

    // zero the rings the warm-up may read before writing
    for (int i = lane; i < 256; i += 32) { m2r[i] = 0.f; bbr[i] = 0.f; }
    for (int i = lane; i < C::PRR * 128; i += 32) prr[i] = 0.f;
    if (lane == 0) {
        for (int s = 0; s < C::NST; s++) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
        const int pre = nstages < C::NST ? nstages : C::NST;
        for (int s = 0; s < pre; s++) {
            mbar_expect_tx(bar0 + 8 * s, 2048);
            tma_bulk_g2s(iq_s + 2048 * s, seg_ptr(a.S, (g0 + s) * AMB_STAGE), 2048, bar0 + 8 * s);
        }
    }

    const float cT = a.P.cT, oe = a.P.one_eps, cTg = a.P.cT * a.P.gfac;
    const int po1 = a.P.po1, po2 = a.P.po2, po3 = a.P.po3;
    const int j_lo = a.j_lo, j_hi = a.j_hi;
    constexpr bool kHasHi = (C::LMOD != 0);
    const bool hi = kHasHi && (4 * lane < C::LMOD);
    const int rows_back = C::RB + (hi ? 1 : 0);
    const int lq = (lane - (C::L / 4)) & 31;

    float rth[C::RB + 1];
#pragma unroll
    for (int m = 0; m <= C::RB; m++) rth[m] = 0.f;
    float pb0 = 0.f, pb1 = 0.f, pb2 = 0.f, pb3 = 0.f;      // previous row: bbs quad
    float pt0 = 0.f, pt1 = 0.f, pt2 = 0.f, pt3 = 0.f;      // previous row: lowered thresholds
    uint32_t cw = 0, cnt = 0;

    for (int gl = 0; gl < nstages; gl++) {
        const int slot = gl & (C::NST - 1);
        const uint32_t parity = (gl / C::NST) & 1;
        while (!mbar_try_wait(bar0 + 8 * slot, parity)) {}
        const unsigned char* st = ws + slot * 2048;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const int k = ((g0 + gl) << 1) + half;
            if (k > rb) break;
            // ---- |x|^2 of this lane's 4 consecutive samples
            const float4 v0 = *reinterpret_cast<const float4*>(st + half * 1024 + lane * 32);
            const float4 v1 = *reinterpret_cast<const float4*>(st + half * 1024 + lane * 32 + 16);
            float m0 = fmaf(v0.x, v0.x, v0.y * v0.y);
            float m1 = fmaf(v0.z, v0.z, v0.w * v0.w);
            float m2 = fmaf(v1.x, v1.x, v1.y * v1.y);
            float m3 = fmaf(v1.z, v1.z, v1.w * v1.w);
            // ---- pulse matched filter: unscaled sum of the last SPC samples, all-positive adds
            float b0, b1, b2, b3;
            if (FL > 1) {
                const int mb = (k & 1) * 128 + 4 * lane;
                *reinterpret_cast<float4*>(m2r + mb) = make_float4(m0, m1, m2, m3);
                __syncwarp();
                float v[FL + 3];
                v[FL - 1] = m0; v[FL] = m1; v[FL + 1] = m2; v[FL + 2] = m3;
#pragma unroll
                for (int t = 1; t < FL; t++) v[FL - 1 - t] = m2r[(mb - t) & 255];
                b0 = v[FL - 1]; b1 = v[FL]; b2 = v[FL + 1]; b3 = v[FL + 2];
#pragma unroll
                for (int t = 1; t < FL; t++) {
                    b0 += v[FL - 1 - t]; b1 += v[FL - t]; b2 += v[FL + 1 - t]; b3 += v[FL + 2 - t];
                }
            } else {
                b0 = m0; b1 = m1; b2 = m2; b3 = m3;
            }
            // ---- row-local inclusive prefix (warp scan of quad totals)
            const float q1 = b0 + b1, q2 = q1 + b2, q3 = q2 + b3;
            float inc = q3;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float y = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += y;
            }
            float exc = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) exc = 0.f;
            const float Rt = __shfl_sync(FULL, inc, 31);
            const float p0 = exc + b0, p1 = exc + q1, p2 = exc + q2, p3 = exc + q3;
            *reinterpret_cast<float4*>(bbr + (k & 1) * 128 + 4 * lane) = make_float4(b0, b1, b2, b3);
            *reinterpret_cast<float4*>(prr + (k & (C::PRR - 1)) * 128 + 4 * lane) = make_float4(p0, p1, p2, p3);
            __syncwarp();
            // ---- noise-floor window and lowered thresholds of row k
            float A_lo = 0.f;
#pragma unroll
            for (int m = 0; m < C::RB; m++) A_lo += rth[m];
            const float A = hi ? (A_lo + rth[C::RB]) : A_lo;
            const float4 pd = *reinterpret_cast<const float4*>(prr + ((k - rows_back) & (C::PRR - 1)) * 128 + 4 * lq);
            const float g = cTg * (Rt + A);
            const float t0 = fmaf(cT, (A - pd.x) + p0, -g);
            const float t1 = fmaf(cT, (A - pd.y) + p1, -g);
            const float t2 = fmaf(cT, (A - pd.z) + p2, -g);
            const float t3 = fmaf(cT, (A - pd.w) + p3, -g);
            // ---- evaluate row k-1 (its look-ahead reaches into row k, now in the ring)
            const int ke = k - 1;
            if (ke >= ra) {
                float nx3 = __shfl_down_sync(FULL, pb0, 1);
                const float c0 = __shfl_sync(FULL, b0, 0);
                if (lane == 31) nx3 = c0;
                const int jb = ke * AMB_ROW + 4 * lane;
                uint32_t m = 0;
                if (pb0 >= pt0 && pb1 <= pb0 * oe && jb >= j_lo && jb < j_hi) m |= 1u;
                if (pb1 >= pt1 && pb2 <= pb1 * oe && jb + 1 >= j_lo && jb + 1 < j_hi) m |= 2u;
                if (pb2 >= pt2 && pb3 <= pb2 * oe && jb + 2 >= j_lo && jb + 2 < j_hi) m |= 4u;
                if (pb3 >= pt3 && nx3 <= pb3 * oe && jb + 3 >= j_lo && jb + 3 < j_hi) m |= 8u;
                if (m) {
                    const int rbase = (ke & 1) * 128 + 4 * lane;
                    const float th[4] = {pt0, pt1, pt2, pt3};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        if (m & (1u << r)) {
                            const float x1 = bbr[(rbase + r + po1) & 255];
                            const float x2 = bbr[(rbase + r + po2) & 255];
                            const float x3 = bbr[(rbase + r + po3) & 255];
                            if (!(fminf(fminf(x1, x2), x3) >= th[r])) m &= ~(1u << r);
                        }
                    }
                }
                const uint32_t e0 = __ballot_sync(FULL, m & 1u), e1 = __ballot_sync(FULL, m & 2u);
                const uint32_t e2 = __ballot_sync(FULL, m & 4u), e3 = __ballot_sync(FULL, m & 8u);
                if (e0 | e1 | e2 | e3) {
                    if (lane < 4) {   // natural bit order: word q covers samples 32q..32q+31 of the row
                        const int sh = 8 * lane;
                        const uint32_t w = spread8((e0 >> sh) & 0xffu) | (spread8((e1 >> sh) & 0xffu) << 1) |
                                           (spread8((e2 >> sh) & 0xffu) << 2) | (spread8((e3 >> sh) & 0xffu) << 3);
                        a.fine[(size_t)ke * 4 + lane] = w;
                    }
                    cnt += __popc(e0) + __popc(e1) + __popc(e2) + __popc(e3);
                    cw |= 1u << (ke & 31);
                }
                if ((ke & 31) == 31 || ke == rb - 1) {
                    if (lane == 0) a.coarse[ke >> 5] = cw;
                    cw = 0;
                }
            }
            // ---- rotate
#pragma unroll
            for (int m = C::RB; m > 0; m--) rth[m] = rth[m - 1];
            rth[0] = Rt;
            pb0 = b0; pb1 = b1; pb2 = b2; pb3 = b3;
            pt0 = t0; pt1 = t1; pt2 = t2; pt3 = t3;
            __syncwarp();
        }
        // ---- refill this slot with stage gl+NST
        if (lane == 0 && gl + C::NST < nstages) {
            fence_proxy_async();
            mbar_expect_tx(bar0 + 8 * slot, 2048);
            tma_bulk_g2s(iq_s + 2048 * slot, seg_ptr(a.S, (g0 + gl + C::NST) * AMB_STAGE), 2048, bar0 + 8 * slot);
        }
    }
    if (lane == 0) a.span_count[span] = cnt;
}

size_t amb_scan_smem_bytes(int spc_i)
{
    switch (spc_i) {
#define CASE(n) case n: return 4 * (size_t)ScanCfg<n, true>::WARP_BYTES;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return 0;
}

template <int SPC, bool PMF>
static cudaError_t launch_scan_t(const AmbScanArgs& a, cudaStream_t s)
{
    const size_t smem = 4 * (size_t)ScanCfg<SPC, PMF>::WARP_BYTES;
    cudaError_t e = cudaFuncSetAttribute(amb_scan_kernel<SPC, PMF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int blocks = (a.n_spans + 3) / 4;
    amb_scan_kernel<SPC, PMF><<<blocks, 128, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t amb_launch_scan(const AmbScanArgs& a, int, cudaStream_t s)
{
    switch (a.P.spc_i) {
#define CASE(n) case n: return a.P.use_pmf ? launch_scan_t<n, true>(a, s) : launch_scan_t<n, false>(a, s);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: break;
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// compaction: (coarse, fine) bitmap -> ordered candidate list. One warp per scan span, so the order is
// span order x row order x bit order = ascending sample index. Offsets come from the per-span counts.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) amb_compact_kernel(const AmbScanArgs a, int* __restrict__ cand_j,
                                                          unsigned int cap, AmbCounters* ctr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span = blockIdx.x * 4 + warp;
    if (span >= a.n_spans) return;
    unsigned int off = 0;
    for (int w = lane; w < span; w += 32) off += a.span_count[w];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) off += __shfl_xor_sync(FULL, off, d);
    const int ra = a.row_lo + span * a.rows_per_span;
    const int rb = min(ra + a.rows_per_span, a.row_hi);
    const int cw_end = (rb + 31) >> 5;
    unsigned int running = off;
    for (int cwb = ra >> 5; cwb < cw_end; cwb += 32) {
        const int idx = cwb + lane;
        const uint32_t cw = idx < cw_end ? a.coarse[idx] : 0u;
        unsigned int cnt = 0;
        for (uint32_t m = cw; m; m &= m - 1) {
            const int row = idx * 32 + (__ffs(m) - 1);
            const uint4 f = *reinterpret_cast<const uint4*>(a.fine + (size_t)row * 4);
            cnt += __popc(f.x) + __popc(f.y) + __popc(f.z) + __popc(f.w);
        }
        unsigned int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int y = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += y;
        }
        unsigned int o = running + inc - cnt;
        for (uint32_t m = cw; m; m &= m - 1) {
            const int row = idx * 32 + (__ffs(m) - 1);
            const uint4 f = *reinterpret_cast<const uint4*>(a.fine + (size_t)row * 4);
            const uint32_t fw[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int q = 0; q < 4; q++)
                for (uint32_t w = fw[q]; w; w &= w - 1) {
                    if (o < cap) cand_j[o] = row * AMB_ROW + q * 32 + (__ffs(w) - 1);
                    o++;
                }
        }
        running += __shfl_sync(FULL, inc, 31);
    }
    if (span == a.n_spans - 1 && lane == 0) {
        ctr->ncand = running < cap ? running : cap;
        if (running > cap) ctr->overflow = 1;
    }
}

cudaError_t amb_launch_compact(const AmbScanArgs& a, int* cand_j, unsigned int cand_cap, AmbCounters* ctr, cudaStream_t s)
{
    amb_compact_kernel<<<(a.n_spans + 3) / 4, 128, 0, s>>>(a, cand_j, cand_cap, ctr);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// canonical arithmetic (the definition DESIGN.md gives; the CPU checker used by tests/ states the same):
//   m2  = fl(fl(re*re) + fl(im*im))                         complex_to_mag_squared, no FMA
//   bb  = fl( (float)(fp64 ascending sum of spc_i m2) * scale_p )   moving_average_ff(spc, 1/spc)   rx_path.py:49
//   avg = fl( (float)(fp64 ascending sum of L bb)   * scale_a )     moving_average_ff(48 spc, ...)  rx_path.py:54
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float canon_m2(const AmbSegs& S, int j)
{
    if (j < 0 || j >= S.n_carry + S.n_main + S.n_tail) return 0.f;
    const float2 v = *seg_ptr(S, j);
    return __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y));
}
// split form: the caller supplies the two float streams; x is in reported coordinates (history-biased)
__device__ __forceinline__ float stream_at(const float* in, long long n, int H, long long x)
{
    const long long k = x - H;
    return (k >= 0 && k < n) ? in[k] : 0.f;
}

#define EX_MAXB 704   // >= L + maxlate + fwd + spc_i for spc_i <= 10

// One warp per candidate. Writes info = late | real<<8 | valid<<9 and avg at the shifted index.
template <bool STREAMS>
__global__ void __launch_bounds__(128) amb_exact_kernel(const AmbExactArgs a)
{
    __shared__ float s_m2[4][EX_MAXB];
    __shared__ float s_bb[4][EX_MAXB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* m2s = s_m2[warp];
    float* bbs = s_bb[warp];
    const AmbParams& P = a.P;
    const int L = P.L, spc = P.spc_i, maxlate = P.maxlate;
    const int fl = P.use_pmf ? spc : 1;
    const int NB = STREAMS ? (maxlate + P.fwd + 1) : (L + maxlate + P.fwd + 1);
    const int NM = NB + fl - 1;
    const int c0off = STREAMS ? 0 : (L - 1);       // bbs index of the candidate start
    const unsigned int ncand = a.ctr->ncand;
    const int nwarps = gridDim.x * 4;
    for (unsigned int ci = blockIdx.x * 4 + warp; ci < ncand; ci += nwarps) {
        const int c = a.cand_j[ci];
        float avgk = 0.f;
        if (STREAMS) {
            for (int i = lane; i < NB; i += 32) bbs[i] = stream_at(a.in0, a.n_streams, P.H, (long long)c + i);
            if (lane <= maxlate) avgk = stream_at(a.in1, a.n_streams, P.H, (long long)c + lane);
            __syncwarp();
        } else {
            const int b_bb = c - L + 1;               // bbs[i] <-> bb[b_bb + i]
            const int b_m2 = b_bb - (fl - 1);
            for (int i = lane; i < NM; i += 32) m2s[i] = canon_m2(a.S, b_m2 + i);
            __syncwarp();
            for (int i = lane; i < NB; i += 32) {
                if (P.use_pmf) {
                    double acc = 0.0;
                    for (int t = 0; t < fl; t++) acc += (double)m2s[i + t];
                    bbs[i] = __fmul_rn((float)acc, P.scale_p);
                } else {
                    bbs[i] = m2s[i];
                }
            }
            __syncwarp();
            if (lane <= maxlate) {                    // inavg at c+lane: window bb[c+lane-L+1 .. c+lane]
                double acc = 0.0;
                for (int t = 0; t < L; t++) acc += (double)bbs[lane + t];
                avgk = __fmul_rn((float)acc, P.scale_a);
            }
        }
        // correlate_preamble (preamble_impl.cc:88-98) at c+k, k = lane-16 in [0, maxlate+1]
        double corrk = 0.0;
        if (lane >= 16 && lane - 16 <= maxlate + 1) {
            const float* q = bbs + c0off + (lane - 16);
            for (int t = 0; t < spc; t++) corrk += (double)q[t];
            for (int t = 0; t < spc; t++) corrk += (double)q[2 * spc + t];
            for (int t = 0; t < spc; t++) corrk += (double)q[7 * spc + t];
            for (int t = 0; t < spc; t++) corrk += (double)q[9 * spc + t];
        }
        const float* in = bbs + c0off;                // in[x] == reference in[i+x] at the candidate start
        const float avg0 = __shfl_sync(FULL, avgk, 0);
        const float pulse_threshold = __fmul_rn(avg0, P.thr);                       // :173
        bool real = in[0] > pulse_threshold;                                        // :174
        if (real && (in[1] > in[0])) real = false;                                  // :175
        if (real && (in[P.po1] < pulse_threshold)) real = false;                    // :177
        if (real && (in[P.po2] < pulse_threshold)) real = false;                    // :178
        if (real && (in[P.po3] < pulse_threshold)) real = false;                    // :179
        uint32_t info = 0;
        float avg_fin = avg0;
        if (real) {
            int i = 0, how_late = 0;
            bool late;
            do {                                                                    // :184-192
                const double now_corr = __shfl_sync(FULL, corrk, 16 + i);
                const double late_corr = __shfl_sync(FULL, corrk, 16 + i + 1);
                late = late_corr > now_corr;
                if (late) { i++; how_late++; }
            } while (late && (float)how_late < P.spc_f);
            avg_fin = __shfl_sync(FULL, avgk, i);
            const float* s = in + i;
            const float sum4 = __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[P.po1]), s[P.po2]), s[P.po3]);
            const float avgpeak = (float)((double)sum4 / 4.0);                      // :198-201
            const float space_threshold =
                __fadd_rn(avg_fin, __fdiv_rn(__fsub_rn(avgpeak, avg_fin), P.thr));  // :203
            bool viol = false;
            for (int j = P.qa0 + lane; j <= P.qa1; j += 32) viol |= (s[j] > space_threshold);  // :205-206
            for (int j = P.qb0 + lane; j <= P.qb1; j += 32) viol |= (s[j] > space_threshold);  // :207-208
            const bool valid = !__any_sync(FULL, viol);
            info = (uint32_t)i | (1u << 8) | (valid ? (1u << 9) : 0u);
        }
        if (lane == 0) { a.cand_info[ci] = info; a.cand_avg[ci] = avg_fin; }
        __syncwarp();
    }
}

cudaError_t amb_launch_exact(const AmbExactArgs& a, int sm_count, cudaStream_t s)
{
    const int blocks = sm_count * 8;
    if (a.in0) amb_exact_kernel<true><<<blocks, 128, 0, s>>>(a);
    else amb_exact_kernel<false><<<blocks, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// resolver: exact sequential restatement of the scan loop's control flow over the candidate list.
// State (pos, p) persists across calls in AmbWalkState. See DESIGN.md "visit order".
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ninputs_of(long long ntot, long long pos, int spc_i)
{
    long long R = ntot - pos;                       // items the scheduler stand-in hands over
    if (R > 0x7fffffffLL) R = 0x7fffffffLL;
    long long n = R - (R % spc_i) - spc_i;          // preamble_impl.cc:150
    return n > 0 ? n : 0;
}

__global__ void amb_walk_seq_kernel(const AmbWalkArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const AmbParams& P = a.P;
    AmbWalkState st = *a.st;
    const int n = (int)a.ctr->ncand;
    unsigned int ndet = 0, nreal = 0;
    for (int k = 0; k < n; k++) if (a.cand_info[k] & (1u << 8)) nreal++;
    int idx = 0;
    while (!st.done) {
        long long ninputs = 0, limit;
        if (a.flush) {
            ninputs = ninputs_of(a.ntot, st.pos, P.spc_i);
            if (ninputs <= 0) { st.done = 1; break; }            // :151 consume_each(0); return 0
            limit = st.pos + ninputs;                            // loop bound i < ninputs (:172)
        } else {
            limit = a.r_safe;
        }
        // first real candidate with start >= p (p can step back by one after a "no room" retry)
        while (idx > 0 && a.org + a.cand_j[idx - 1] >= st.p) idx--;
        while (idx < n && (!(a.cand_info[idx] & (1u << 8)) || a.org + a.cand_j[idx] < st.p)) idx++;
        if (idx >= n || a.org + a.cand_j[idx] >= limit) {
            if (a.flush) { st.pos += ninputs; st.p = st.pos; continue; }   // :244 consume_each(ninputs)
            if (st.p < limit) st.p = limit;                      // everything before r_safe has been looked at
            break;
        }
        const uint32_t info = a.cand_info[idx];
        const long long fin = a.org + a.cand_j[idx] + (long long)(info & 0xffu);
        if (!(info & (1u << 9))) { st.p = fin + 1; idx++; continue; }      // :209 continue -> i++
        const long long i_rel = fin - st.pos;
        if (a.flush && (float)(ninputs - i_rel) < P.skip_f) {              // :212 no room
            const long long consumed = i_rel - 1 > 0 ? i_rel - 1 : 0;      // :213
            if (consumed == 0) { st.done = 1; break; }
            st.pos += consumed; st.p = st.pos;
            continue;
        }
        a.cand_info[idx] = info | (1u << 10);                              // accepted: 240-chip packet
        ndet++;
        const long long consumed = (long long)(int)((float)i_rel + P.skip_f);  // :237 float arithmetic
        st.pos += consumed; st.p = st.pos;
        idx++;
    }
    st.ncand_real += nreal; st.ndet += ndet;
    *a.st = st;
    a.ctr->ndet_call = ndet;
    a.ctr->nreal_call = nreal;
}

cudaError_t amb_launch_walk_seq(const AmbWalkArgs& a, cudaStream_t s)
{
    amb_walk_seq_kernel<<<1, 32, 0, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// slicer: one warp per accepted preamble
// ------------------------------------------------------------------------------------------------
// llslicer (slicer_impl.cc:67-100): bit 0 = decision, bit 1 = confidence
__device__ __forceinline__ int llslice(float bit0, float bit1, float ref)
{
    const float highlimit = (float)((double)ref * 1.414);     // :71
    const float lowlimit = (float)((double)ref * 0.707);      // :72
    const bool f = (bit0 > lowlimit) && (bit0 < highlimit);
    const bool s = (bit1 > lowlimit) && (bit1 < highlimit);
    if (f && !s) return 1 | 2;
    if (s && !f) return 0 | 2;
    if (f && s) return (bit0 > bit1) ? 1 : 0;
    const bool d = bit0 > bit1;
    bool c;
    if (d) c = (double)bit1 < (double)lowlimit * 0.5;         // :91
    else c = (double)bit0 < (double)lowlimit * 0.5;           // :94
    return (d ? 1 : 0) | (c ? 2 : 0);
}

// Packet rules of slicer_impl::work (slicer_impl.cc:117-182) on 240 chips held in shared memory.
// Executed by a full warp; lane 0 fills *f (sample_index/secs/frac are the caller's business).
__device__ void slice_packet_warp(const float* chips, amb_frame* f, int lane)
{
    const float ref = (float)((double)__fadd_rn(__fadd_rn(__fadd_rn(chips[0], chips[2]), chips[7]), chips[9]) / 4.0); // :128-131
    uint32_t dw[4], lw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {                      // bit j = 32k + lane, chips 16+2j, 17+2j (:133,:147)
        const int j = 32 * k + lane;
        int r = 2;
        if (j < 112) r = llslice(chips[16 + 2 * j], chips[17 + 2 * j], ref);
        dw[k] = __brev(__ballot_sync(FULL, r & 1));    // MSB-first: bit 31 <-> j = 32k
        lw[k] = __ballot_sync(FULL, !(r & 2));         // bit lane <-> low confidence at j
    }
    const unsigned hdr = dw[0] >> 27;                                         // :135-139
    const bool is_long = (hdr == 16 || hdr == 17 || hdr == 20 || hdr == 21); // :140
    const int nbits = is_long ? 112 : 56;                                     // :142
    if (!is_long) { dw[1] &= 0xFFFFFF00u; dw[2] = 0; dw[3] = 0; lw[1] &= 0x00FFFFFFu; lw[2] = 0; lw[3] = 0; }
    else { dw[3] &= 0xFFFF0000u; lw[3] &= 0x0000FFFFu; }
    // CRC over the first nbits-24 bits (modes_crc.cc:55-63) as an XOR fold of per-bit remainders
    uint32_t crc = 0;
    const int nmsg = nbits - 24;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int j = 32 * k + lane;
        if (j < nmsg && ((dw[k] >> (31 - lane)) & 1u)) crc ^= c_crc_rem[nmsg - 1 - j];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(FULL, crc, d);
    if (lane == 0) {
        unsigned char data[14];
#pragma unroll
        for (int m = 0; m < 14; m++) data[m] = (unsigned char)(dw[m >> 2] >> (24 - 8 * (m & 3)));
        unsigned numlow = 0;
#pragma unroll
        for (int m = 0; m < 24; m++) f->lowconfbits[m] = 0;
        for (int k = 0; k < 4 && numlow < 24; k++)                            // :152-158, ascending j, cap 24
            for (uint32_t w = lw[k]; w && numlow < 24; w &= w - 1) f->lowconfbits[numlow++] = (uint8_t)(32 * k + __ffs(w) - 1);
        const unsigned df = (data[0] >> 3) & 0x1F;                            // :168
        const uint32_t ap = ((uint32_t)data[nbits / 8 - 3] << 16) | ((uint32_t)data[nbits / 8 - 2] << 8) | data[nbits / 8 - 1];
        bool zeroes = true;                                                   // :162-166
#pragma unroll
        for (int m = 0; m < 14; m++) if (data[m]) zeroes = false;
        bool passed = !zeroes;
        if (passed && !is_long && df != 11 && numlow > 0) passed = false;     // :170
        if (passed && df == 11 && numlow >= 10) passed = false;               // :171
        uint32_t syn = 0;
        if (passed) {
            syn = crc ^ ap;                                                   // :173-177
            if (syn && (df == 11 || df == 17)) passed = false;                // :182
        }
#pragma unroll
        for (int m = 0; m < 14; m++) f->data[m] = data[m];
        f->ref_level = ref;
        f->crc = syn;
        f->nbits = (uint8_t)nbits;
        f->df = (uint8_t)df;
        f->numlowconf = (uint8_t)numlow;
        f->passed = passed ? 1 : 0;
    }
}

template <bool STREAMS>
__global__ void __launch_bounds__(128) amb_slice_kernel(const AmbSliceArgs a)
{
    __shared__ float s_chips[4][240];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* chips = s_chips[warp];
    const AmbParams& P = a.P;
    const unsigned int ncand = a.ctr->ncand;
    const int nwarps = gridDim.x * 4;
    const int fl = P.use_pmf ? P.spc_i : 1;
    for (unsigned int ci = blockIdx.x * 4 + warp; ci < ncand; ci += nwarps) {
        const uint32_t info = a.cand_info[ci];
        if (!(info & (1u << 10))) continue;                    // warp-uniform
        const int fin = a.cand_j[ci] + (int)(info & 0xffu);
        const float avg_fin = a.cand_avg[ci];
        for (int j = lane; j < 240; j += 32) {                 // preamble_impl.cc:219-221
            const int n = fin + c_chip_off[j];
            float bb;
            if (STREAMS) {
                bb = stream_at(a.in0, a.n_streams, P.H, (long long)n);
            } else if (P.use_pmf) {
                double acc = 0.0;
                for (int t = 0; t < fl; t++) acc += (double)canon_m2(a.S, n - fl + 1 + t);
                bb = __fmul_rn((float)acc, P.scale_p);
            } else {
                bb = canon_m2(a.S, n);
            }
            chips[j] = __fsub_rn(bb, avg_fin);
        }
        __syncwarp();
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(&a.ctr->nframes, 1u);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot < a.frame_cap) {
            amb_frame* f = a.frames + slot;
            slice_packet_warp(chips, f, lane);
            if (lane == 0) {
                f->sample_index = (uint64_t)(a.org + fin);
                f->secs = 0; f->frac = 0.0;
                for (int m = 0; m < 6; m++) f->pad_[m] = 0;
                if (f->passed) atomicAdd(&a.ctr->npassed_call, 1u);
            }
            if (a.chips_out) for (int j = lane; j < 240; j += 32) a.chips_out[(size_t)slot * 240 + j] = chips[j];
        } else if (lane == 0) {
            a.ctr->frame_overflow = 1;
        }
        __syncwarp();
    }
}

cudaError_t amb_launch_slice(const AmbSliceArgs& a, int sm_count, cudaStream_t s)
{
    const int blocks = sm_count * 4;
    if (a.in0) amb_slice_kernel<true><<<blocks, 128, 0, s>>>(a);
    else amb_slice_kernel<false><<<blocks, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// slicer only (split-form block): packets of 240 chips already in device memory
__global__ void __launch_bounds__(128) amb_slice_chips_kernel(const float* __restrict__ chips_in, int ndet, amb_frame* frames)
{
    __shared__ float s_chips[4][240];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* chips = s_chips[warp];
    for (int d = blockIdx.x * 4 + warp; d < ndet; d += gridDim.x * 4) {
        for (int j = lane; j < 240; j += 32) chips[j] = chips_in[(size_t)d * 240 + j];
        __syncwarp();
        slice_packet_warp(chips, frames + d, lane);
        __syncwarp();
    }
}
cudaError_t amb_launch_slice_chips(const float* chips, int ndet, amb_frame* frames, cudaStream_t s)
{
    if (ndet <= 0) return cudaSuccess;
    int blocks = (ndet + 3) / 4; if (blocks > 2048) blocks = 2048;
    amb_slice_chips_kernel<<<blocks, 128, 0, s>>>(chips, ndet, frames);
    return cudaGetLastError();
}

// device CRC parity hook: same XOR fold as the slicer
__global__ void amb_crc_kernel(const uint8_t* __restrict__ data, int n, int length, uint32_t* out)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const int nmsg = length * 8;
    uint32_t crc = 0;
    for (int j = lane; j < nmsg; j += 32)
        if ((data[(size_t)w * length + (j >> 3)] >> (7 - (j & 7))) & 1) crc ^= c_crc_rem[nmsg - 1 - j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(FULL, crc, d);
    if (lane == 0) out[w] = crc;
}
cudaError_t amb_launch_crc(const uint8_t* data, int n, int length, uint32_t* out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    amb_crc_kernel<<<(n + 3) / 4, 128, 0, s>>>(data, n, length, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// carry: keep the last kc samples of (carry ++ new data) for the next call
// ------------------------------------------------------------------------------------------------
__global__ void amb_carry_kernel(const AmbSegs S, float2* __restrict__ dst, int kc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kc) return;
    dst[i] = *seg_ptr(S, S.n_valid - kc + i);
}
cudaError_t amb_launch_carry(const AmbSegs& S, float2* dst, int kc, cudaStream_t s)
{
    amb_carry_kernel<<<(kc + 255) / 256, 256, 0, s>>>(S, dst, kc);
    return cudaGetLastError();
}

cudaError_t amb_upload_tables(const int* chip_off)
{
    cudaError_t e = cudaMemcpyToSymbol(c_chip_off, chip_off, 240 * sizeof(int));
    if (e != cudaSuccess) return e;
    unsigned int rem[96];
    unsigned int r = 0xFFF409u;                         // x^24 mod G
    for (int t = 0; t < 96; t++) {
        rem[t] = r;
        r = (r & 0x800000u) ? (((r << 1) ^ 0xFFF409u) & 0xFFFFFFu) : ((r << 1) & 0xFFFFFFu);
    }
    return cudaMemcpyToSymbol(c_crc_rem, rem, sizeof(rem));
}
