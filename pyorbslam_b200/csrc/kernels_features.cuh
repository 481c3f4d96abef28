// K4+K6: orientation (IC_Angle) + rBRIEF descriptor + output assembly; K7+K8: stereo matching.
// Float arithmetic here must reproduce the host bit-for-bit: the translation unit is compiled with
// --fmad=false (no contraction) and divisions use the IEEE-rounded intrinsics.
#pragma once
#include "plan.h"
#include "../../include/b200orb_pattern31.h"

// The 256 test pairs (x0, y0, x1, y1; include/b200orb_pattern31.h) reach the descriptor kernel as floats in a host-built table
// (b200orb.cu), transposed so that a warp's read of "pair k of every lane" is one contiguous 512-byte line:
// fpat[k * 32 + lane] = pair 8 * lane + k.  Lane i produces descriptor byte i.

// cv::fastAtan2 (degrees), scalar polynomial path
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = __fdiv_rn(ax, ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// glibc 2.39 sinf/cosf (ARM optimized-routines sincosf) on [0, 2*pi]: fp64 polynomial, one rounding to fp32
// (the CPU twin in oracle/cvprims.hpp is checked against the host libm on every float).
__device__ __forceinline__ double sc_sin_poly(double x, double x2) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double x3 = x * x2, s1p = s2 + x2 * s3, x7 = x3 * x2, s = x + x3 * s1;
    return s + x7 * s1p;
}
__device__ __forceinline__ double sc_cos_poly(double x2, bool neg) {
    double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
           c4 = 0x1.99343027bf8c3p-16;
    if (neg) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; c4 = -c4; }
    const double x4 = x2 * x2, c2p = c3 + x2 * c4, c1p = c1 + x2 * c2, x6 = x4 * x2, c = c0 + x2 * c1p;
    return c + x6 * c2p;
}
__device__ __forceinline__ void glibc_sincosf(float y, float* sp, float* cp) {
    double x = (double)y;
    const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
    if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff)) {
        if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ff)) { *sp = y; *cp = 1.0f; return; }
        const double x2 = x * x;
        *sp = (float)sc_sin_poly(x, x2);
        *cp = (float)sc_cos_poly(x2, false);
        return;
    }
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = ((int)r + 0x800000) >> 24;
    x = x - (double)n * 0x1.921FB54442D18p0;
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool tab = (n & 2) != 0;
    const double xs = x * s, x2 = x * x;
    *sp = (float)((n & 1) ? sc_cos_poly(x2, tab) : sc_sin_poly(xs, x2));
    *cp = (float)(((n ^ 1) & 1) ? sc_cos_poly(x2, tab) : sc_sin_poly(xs, x2));
}

// ------------------------------------------------------------------------------------------------
// K4 + K6 + a10: one warp per output keypoint.  Level keypoints come from K3 in list order; the warp finds
// its level by the per-level counts (levels ascending, ORBextractor.cpp:1075-1103), computes
//   * IC_Angle (:77-104): m10/m01 over the radius-15 disc on the UN-blurred level, lanes = columns,
//     31 coalesced row reads, warp-shuffle reduction, fastAtan2;
//   * computeOrbDescriptor (:108-147): lane i produces descriptor byte i from 16 rotated samples of the
//     BLURRED level, coordinates rounded half-to-even, a = cosf, b = sinf of angle * (float)(pi/180);
//   * the output row: (x, y) * mvScaleFactor[level] for level > 0 (:1094-1100), size, angle, response, octave.
// ------------------------------------------------------------------------------------------------
#define DESC_WARPS 8
#define DESC_KPW 8     // keypoints per warp
// pitch of the staged descriptor window = width of its TMA box: 37 columns + up to 15 bytes of alignment shift need 52; 80 rather
// than 64 because with 16-word rows the samples of a warp (clustered around the window centre, ~10 words wide) only ever touch ~20 of the
// 32 banks (ncu: 4.75 wavefronts per sample load); 20-word rows walk through all banks with a period of 8 rows
#define DESC_PP 80
#define DESC_WIN (24 * 128)   // bytes per warp: 37 rows x 80, rounded up to the 128-byte alignment a TMA destination needs
__device__ __forceinline__ int dp4a_us(u32 a_u8x4, u32 b_s8x4, int c) {     // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__global__ void __launch_bounds__(DESC_WARPS * 32, 6) k_describe(const __grid_constant__ Plan P, const __grid_constant__ LevelMaps BM,
                                                              const u8* __restrict__ pyr, const u32* __restrict__ lvl_kp,
                                                              const int* __restrict__ lvl_cnt, const uint2* __restrict__ mtab,
                                                              const float4* __restrict__ fpat, float* __restrict__ kps,
                                                              u8* __restrict__ desc, int* __restrict__ nkp) {
    pdl_enter();
    const int slot = blockIdx.y, lane = threadIdx.x & 31;
    // the moment table goes to shared memory once per CTA, the float pattern into registers once per warp; every warp then
    // walks DESC_KPW consecutive keypoints
    __shared__ __align__(128) u8 s_patch[DESC_WARPS * DESC_WIN];   // per warp: 37 rows x DESC_PP bytes of the blurred level (one TMA box)
    __shared__ unsigned long long s_bar[DESC_WARPS];
    __shared__ float4 s_fpat[256];
    for (int k = threadIdx.x; k < 256; k += DESC_WARPS * 32) s_fpat[k] = __ldg(fpat + k);
    // level of keypoint i: lanes hold the running ends of the per-level counts
    int lend = lane < P.nlevels ? lvl_cnt[(size_t)slot * P.nlevels + lane] : 0;
#pragma unroll
    for (int d = 1; d < ORB_MAX_LEVELS; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, lend, d);
        if (lane >= d) lend += up;
    }
    const int total = __shfl_sync(0xffffffffu, lend, ORB_MAX_LEVELS - 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) nkp[slot] = total;
    __syncthreads();
    const int i0 = (blockIdx.x * DESC_WARPS + (threadIdx.x >> 5)) * DESC_KPW;
    u8* const win = s_patch + (threadIdx.x >> 5) * DESC_WIN;
    unsigned long long* const bar = &s_bar[threadIdx.x >> 5];
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    const CUtensorMap* const bmaps = BM.m;
    u32 parity = 0;
    for (int i = i0; i < min(i0 + DESC_KPW, total); ++i) {
    const int l = __popc(__ballot_sync(0xffffffffu, lane < P.nlevels && i >= lend));
    const int off = __shfl_sync(0xffffffffu, lend, max(l - 1, 0)) & (l ? -1 : 0);
    const LevelGeom& G = P.lv[l];
    const u32 packed = lvl_kp[(size_t)slot * P.kp_total + G.kp_ofs + (i - off)];
    const int x = packed & 0xfff, y = (packed >> 12) & 0xfff, resp = packed >> 24;

    // The descriptor stage below samples a 37 x 37 window of the blurred level: one TMA box copy (64 x 37 bytes from the 16-byte
    // boundary below the window's left edge; rows / columns outside the level are zero-filled and never sampled) is issued now and
    // lands in the warp's buffer while the orientation is computed -- no load / store instructions are spent on the staging.
    const int wx0 = (x - 18) & ~15, wshift = (x - 18) - wx0;
    __syncwarp();                                               // the previous keypoint's samples are done
    if (lane == 0) {
        mbar_expect_tx(bar, (u32)(DESC_PP * 37));
        tma_box_g2s(win, bmaps + l, wx0, y - 18, slot, bar);
    }
    // ---- orientation ----
    const u8* c = pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs + (size_t)(y + ORB_EDGE) * G.pitch + (x + ORB_EDGE);
    // m10 = sum u * I, m01 = sum v * I over the radius-15 disc (exact integers, any order): the disc is read as aligned words,
    // 3 rows x 9 words per step, and each word is two 4-way dot products against the step's coefficient words (host table:
    // signed u and signed v of the word's 4 bytes, 0 outside the disc).  Lanes 27..31 read in-bounds pixels against zeros.
    int m10 = 0, m01 = 0;
    {
        const u8* c0 = c - 15;                               // u = -15 on the centre row
        const int al = (int)(reinterpret_cast<size_t>(c0) & 3);
        const int rsub = (lane * 57) >> 9, k = lane - 9 * rsub;   // lane / 9, lane % 9
        const int wpr = G.pitch >> 2;
        const u32* p = reinterpret_cast<const u32*>(c0 - al) + (ptrdiff_t)((rsub - 15) * wpr + k);
        const uint2* tb = mtab + al * (MOM_STEPS * 32) + lane;          // 11 KB table, read through L1 (in shared memory it cost a CTA per SM)
        const ptrdiff_t step = 3 * wpr;
#pragma unroll
        for (int s = 0; s < MOM_STEPS; ++s, p += step) {
            const u32 w = *p;
            const uint2 cf = __ldg(tb + s * 32);
            m10 = dp4a_us(w, cf.x, m10);
            m01 = dp4a_us(w, cf.y, m01);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- descriptor ----
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    glibc_sincosf(angle * factorPI, &b, &a);
    // Sample coordinates: cvRound(x * b + y * a) rows, cvRound(x * a - y * b) columns (ORBextractor.cpp:117-120, products and sums
    // rounded separately: --fmad=false); |coordinate| <= 18 (the pattern's largest radius is 18.38).
    // The samples come from the window the TMA unit staged in shared memory: read straight from global memory the 16 samples of a
    // lane hit ~25 different cache lines per warp instruction and the L1 tag stage becomes the bound of the kernel.
    // Rounding to nearest-even is done by adding 1.5 * 2^23 (FADD, exact for |v| < 2^22) instead of a float->int conversion: the
    // integer then sits in the low mantissa bits, biased by K = 0x4B400000, and the shared-memory byte address
    // rbits * DESC_PP + cbits + D (mod 2^32) absorbs the bias in the per-keypoint constant D.
    const u32 D = (u32)(18 * DESC_PP + 18 + wshift) - 0x4B400000u * (u32)(DESC_PP + 1);
    mbar_wait(bar, parity);
    parity ^= 1u;
    const float RN = 12582912.f;
    const u8* sp = win;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 q = s_fpat[k * 32 + lane];
        const float r0 = (q.x * b + q.y * a) + RN, c0 = (q.x * a - q.y * b) + RN;
        const float r1 = (q.z * b + q.w * a) + RN, c1 = (q.z * a - q.w * b) + RN;
        const u32 t0 = sp[__float_as_uint(r0) * (u32)DESC_PP + (__float_as_uint(c0) + D)];
        const u32 t1 = sp[__float_as_uint(r1) * (u32)DESC_PP + (__float_as_uint(c1) + D)];
        val |= (t0 < t1) << k;
    }
    desc[((size_t)slot * P.kp_total + i) * 32 + lane] = (u8)val;
    if (lane < 6) {
        float o;
        const float fx = (float)x, fy = (float)y;
        switch (lane) {
            case 0: o = l ? fx * G.sf : fx; break;
            case 1: o = l ? fy * G.sf : fy; break;
            case 2: o = (float)G.psize; break;
            case 3: o = angle; break;
            case 4: o = (float)resp; break;
            default: o = (float)l; break;
        }
        kps[((size_t)slot * P.kp_total + i) * 6 + lane] = o;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// K7a: row index of the RIGHT keypoints -- the GPU form of vRowIndices (Frame.py:170-179), built like the reference builds it:
// a right keypoint is entered into EVERY row of its band floor(y - 2s) .. ceil(y + 2s) (evaluated in double exactly like
// Frame.py:173-176), so a left keypoint at row v reads one bin, rowStart[v] .. rowStart[v + 1], and every entry it finds already
// passed the row test (visiting the +-10 neighbouring bins of a one-entry-per-keypoint index cost 2.3x more candidates, most of them
// low-octave keypoints whose band is only 5 rows).  Counting sort over the bands: histogram, block scan, scatter; 8 bytes per
// entry (uR, octave << 24 | index).  Order inside a bin is irrelevant: the winner is the minimum of (dist << 20 | right index),
// which equals the reference's first strict minimum in ascending index.  One CTA per right image.
// ------------------------------------------------------------------------------------------------
#define RI_THREADS 256
__global__ void __launch_bounds__(RI_THREADS) k_rowindex(const float* __restrict__ kpsR, const int* __restrict__ nR, long long kp_stride,
                                                         int n_stride, int kp_row, int oct_idx, const __grid_constant__ StereoGeom SG,
                                                         int* __restrict__ rowStart, uint2* __restrict__ rmeta,
                                                         long long idx_stride, int* __restrict__ status, int status_stride) {
    pdl_enter();
    const int nRows = SG.nRows;
    extern __shared__ int ri_hist[];     // nRows + 1 counters, then nRows cursors
    __shared__ int ri_tmp[RI_THREADS / 32 + 1];
    const int pair = blockIdx.x;
    status += (size_t)pair * status_stride;          // one flag word per pair in the batch API (stride 0: one word in all)
    if (threadIdx.x == 0) *status = 0;               // cleared here (no memset between the kernels: they are chained programmatically);
    __syncthreads();                                 // k_stereo of the same pair ORs its own flags in afterwards
    const int n = nR[(size_t)pair * n_stride];
    const float* k = kpsR + (size_t)pair * kp_stride;
    int* rs = rowStart + (size_t)pair * (nRows + 1);
    uint2* rm = rmeta + (size_t)pair * idx_stride;
    int* cursor = ri_hist + nRows + 1;
    for (int i = threadIdx.x; i <= nRows; i += RI_THREADS) ri_hist[i] = 0;
    __syncthreads();
    auto band = [&](const float* r, int& minr, int& maxr, int& o) {
        o = min(max((int)r[oct_idx], 0), SG.nlevels - 1);
        const double y = (double)r[1], reach = 2.0 * (double)SG.sf[o];
        minr = (int)floor(y - reach); maxr = (int)ceil(y + reach);
    };
    for (int j = threadIdx.x; j < n; j += RI_THREADS) {
        int minr, maxr, o;
        band(k + (size_t)j * kp_row, minr, maxr, o);
        if (minr < 0 || maxr >= nRows) atomicOr(status, 2);      // vRowIndices[yi] leaves the list (the reference raises / wraps around)
        for (int row = max(minr, 0); row <= min(maxr, nRows - 1); ++row) atomicAdd(&ri_hist[row], 1);
    }
    __syncthreads();
    // exclusive scan over nRows + 1 entries (the last one becomes the total)
    const int len = nRows + 1;
    const int per = (len + RI_THREADS - 1) / RI_THREADS;
    const int b = min((int)threadIdx.x * per, len), e = min(b + per, len);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += ri_hist[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ri_tmp[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += ri_tmp[w];
    base += incl - sum;
    for (int i = b; i < e; ++i) { const int t = ri_hist[i]; ri_hist[i] = base; base += t; }
    __syncthreads();
    for (int i = threadIdx.x; i <= nRows; i += RI_THREADS) { rs[i] = ri_hist[i]; if (i < nRows) cursor[i] = ri_hist[i]; }
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += RI_THREADS) {
        const float* r = k + (size_t)j * kp_row;
        int minr, maxr, o;
        band(r, minr, maxr, o);
        const uint2 ent = make_uint2(__float_as_uint(r[0]), ((unsigned)o << 24) | (unsigned)j);
        for (int row = max(minr, 0); row <= min(maxr, nRows - 1); ++row) rm[atomicAdd(&cursor[row], 1)] = ent;
    }
}

// ------------------------------------------------------------------------------------------------
// K7 + K8: Frame.compute_stereo_matches (Frame.py:161-279), one warp per left keypoint.
//   K7  candidates come from the row index (k_rowindex); per candidate the exact row-band test (right keypoint
//       rows floor(y-2s) .. ceil(y+2s) in double, Frame.py:173-179), octave within +-1, uR in [uL - maxD, uL]
//       (:209-215); Hamming distance = __popc over 2 x 16-byte loads; the reference's "first strict minimum in
//       ascending right index" is the minimum of (dist << 20 | index).
//   K8  11x11 SAD slide (L = 5) on the keypoint's pyramid level through the reference's step-ignoring pyramid
//       view, parabola fit, |delta| > 1 rejection, disparity / depth in float32 exactly as
//       NumPy >= 2 evaluates them.
// kps rows are float[kp_row] with (x, y) first and the octave at index `oct_idx`.
// ------------------------------------------------------------------------------------------------
#define ST_WARPS 8

struct StereoArgs {
    const float* kpsL; const u8* descL; const int* nL;      // per pair: + pair * stride
    const float* kpsR; const u8* descR; const int* nR;
    const u8* pyrL; const u8* pyrR;                         // per pair: + pair * pyr_stride
    long long kp_stride, desc_stride, pyr_stride;           // element / byte strides between pairs
    int n_stride;                                           // stride of nL / nR between pairs (ints)
    int kp_row, oct_idx;                                    // floats per keypoint row, index of the octave
    int out_stride;                                         // rows per pair in the outputs
    const int* rowStart; const uint2* rmeta; long long idx_stride;   // row index of the right keypoints (k_rowindex): bins + entries per pair
    float mbf32, mb, maxD;
    double mbf;
    float* uRight; float* depth; int* matchIdx;
    int* status; int status_stride;                         // range-error flags: status[pair * status_stride]
    int* sadDist;                                           // optional: SAD minimum of accepted matches, -1 otherwise
};

__device__ __forceinline__ const u8* view_ptr(const u8* base, const StereoGeom& SG, int o, int row, int col) {
    const int lin = SG.off0[o] + row * SG.vstride[o] + col;
    int pr = (int)__umulhi((unsigned)lin, SG.magic[o]);           // lin / plog via ceil(2^32 / plog); may overshoot by one
    if (pr * SG.plog[o] > lin) --pr;
    return base + SG.base[o] + (size_t)pr * SG.pitch[o] + (lin - pr * SG.plog[o]);
}

__global__ void __launch_bounds__(ST_WARPS * 32, 8) k_stereo(const __grid_constant__ StereoGeom SG, const StereoArgs A) {
    pdl_enter();
    __shared__ __align__(16) unsigned char s_win[ST_WARPS][11 * 12 + 11 * 24 + 4];     // 400 bytes per warp: word-aligned window rows
    const int pair = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nL = A.nL[(size_t)pair * A.n_stride];
    const int iL = blockIdx.x * ST_WARPS + warp;
    if (iL >= nL) return;                                   // warp-uniform; no block-level barrier below
    const float* kL = A.kpsL + (size_t)pair * A.kp_stride;
    const float* kR = A.kpsR + (size_t)pair * A.kp_stride;
    const u8* dL = A.descL + (size_t)pair * A.desc_stride;
    const u8* dR = A.descR + (size_t)pair * A.desc_stride;
    const int* rs = A.rowStart + (size_t)pair * (SG.nRows + 1);

    const float* p = kL + (size_t)iL * A.kp_row;
    const float uL = p[0], vL = p[1];
    const int oL = (int)p[A.oct_idx];
    const int row = (int)vL;                                // int(vL), Frame.py:192
    unsigned best = (100u << 20);                           // bestDist = TH_HIGH, bestIdxR = 0 (Frame.py:203-204)
    if (row < 0 || row >= SG.nRows) {
        if (lane == 0) atomicOr(A.status + (size_t)pair * A.status_stride, 1);   // the reference indexes vRowIndices[int(vL)] here
    } else if (!(uL < 0)) {                                 // maxU < 0 -> continue (Frame.py:200-201)
        const float minU = uL - A.maxD;
        const uint4* dl = reinterpret_cast<const uint4*>(dL + (size_t)iL * 32);
        const uint4 l0 = dl[0], l1 = dl[1];
        const int c0 = rs[row], c1 = rs[row + 1];                     // vRowIndices[int(vL)], Frame.py:192-194
        const uint2* rm = A.rmeta + (size_t)pair * A.idx_stride;
        for (int c = c0 + lane; c < c1; c += 32) {
            const uint2 m = rm[c];
            const int j = (int)(m.y & 0xffffffu);
            const float uR = __uint_as_float(m.x);
            const int oR = (int)(m.y >> 24);
            if (oR < oL - 1 || oR > oL + 1 || !(minU <= uR) || !(uR <= uL)) continue;      // Frame.py:209-215
            const uint4* dr = reinterpret_cast<const uint4*>(dR + (size_t)j * 32);
            const uint4 r0 = __ldg(dr), r1 = __ldg(dr + 1);
            const unsigned d = __popc(l0.x ^ r0.x) + __popc(l0.y ^ r0.y) + __popc(l0.z ^ r0.z) + __popc(l0.w ^ r0.w) +
                               __popc(l1.x ^ r1.x) + __popc(l1.y ^ r1.y) + __popc(l1.z ^ r1.z) + __popc(l1.w ^ r1.w);
            best = min(best, (d << 20) | (unsigned)j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    }
    {
        const int bestDist = best >> 20, bestR = best & 0xfffff;
        const size_t oi = (size_t)pair * A.out_stride + iL;
        float outU = -1.f, outD = -1.f;
        int outM = -1, outS = -1;
        if (bestDist < 75) {    // thOrbDist = (TH_HIGH + TH_LOW) / 2, Frame.py:166,222
            outM = bestR;
            const float uR0 = kR[(size_t)bestR * A.kp_row];
            const double inv = (double)SG.isf[oL];
            const int su = __double2int_rn((double)uL * inv), sv = __double2int_rn((double)vL * inv);   // python round()
            const int sr = __double2int_rn((double)uR0 * inv);
            const int w = SG.w[oL], h = SG.h[oL];
            bool ok = !(sr < 0 || sr + 11 >= w);                                        // Frame.py:240-243
            if (ok && (sv - 5 < 0 || sv + 5 >= h || su - 5 < 0 || su + 5 >= w || sr - 10 < 0 || sr + 10 >= w)) {
                ok = false;                                                             // the reference would raise
                if (lane == 0) atomicOr(A.status + (size_t)pair * A.status_stride, 1);
            }
            if (ok) {
                // Stage the 11 x 11 left and 11 x 21 right windows of the (sheared) pyramid view.  A window row is a run of consecutive
                // LOGICAL bytes; it is contiguous in the padded buffer too unless it crosses the logical pitch, so lane r (left) /
                // lane 11 + r (right) fetches its whole row as aligned 32-bit words, funnel-shifts them into place and stores words
                // (rows padded to 12 / 24 bytes in shared memory); the rare row that does cross is copied byte by byte.
                unsigned char* wl = s_win[warp];          // [11][12]
                unsigned char* wr = wl + 11 * 12;          // [11][24]
                const u8* bl = A.pyrL + (size_t)pair * A.pyr_stride;
                const u8* br = A.pyrR + (size_t)pair * A.pyr_stride;
                if (lane < 22) {
                    const bool left = lane < 11;
                    const int r = left ? lane : lane - 11, n = left ? 11 : 21;
                    const int lin = SG.off0[oL] + (sv - 5 + r) * SG.vstride[oL] + (left ? su - 5 : sr - 10);
                    int pr = (int)__umulhi((unsigned)lin, SG.magic[oL]);           // lin / plog via ceil(2^32 / plog); may overshoot by one
                    if (pr * SG.plog[oL] > lin) --pr;
                    const int pc = lin - pr * SG.plog[oL];
                    const u8* src = (left ? bl : br) + SG.base[oL] + (size_t)pr * SG.pitch[oL] + pc;
                    unsigned char* dst = left ? wl + r * 12 : wr + r * 24;
                    const size_t addr = reinterpret_cast<size_t>(src);
                    const u32* wp = reinterpret_cast<const u32*>(addr & ~(size_t)3);
                    const int sh = 8 * (int)(addr & 3);
                    u32 v[6];
                    {
                        const u32 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                        v[0] = __funnelshift_r(w0, w1, sh); v[1] = __funnelshift_r(w1, w2, sh); v[2] = __funnelshift_r(w2, w3, sh);
                        if (!left) {
                            const u32 w4 = wp[4], w5 = wp[5], w6 = wp[6];
                            v[3] = __funnelshift_r(w3, w4, sh); v[4] = __funnelshift_r(w4, w5, sh); v[5] = __funnelshift_r(w5, w6, sh);
                        }
                    }
                    const int nfirst = SG.plog[oL] - pc, wrap = SG.pitch[oL] - SG.plog[oL];
                    if (nfirst < n && wrap != 0) {
                        // the row crosses the logical pitch: its bytes from nfirst on continue at the start of the next padded row, i.e.
                        // `wrap` bytes further; fetch that stream the same way and take bytes >= nfirst from it
                        const size_t addr2 = addr + (size_t)wrap;
                        const u32* wq = reinterpret_cast<const u32*>(addr2 & ~(size_t)3);
                        const int sh2 = 8 * (int)(addr2 & 3);
                        const int nw = left ? 3 : 6;
                        u32 prev = wq[0];
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            if (j >= nw) break;
                            const u32 nxt = wq[j + 1];
                            const u32 b = __funnelshift_r(prev, nxt, sh2);
                            prev = nxt;
                            const int keep = min(max(nfirst - 4 * j, 0), 4);              // bytes of word j that come before the wrap
                            const u32 mask = keep >= 4 ? 0xffffffffu : ((1u << (8 * keep)) - 1u);
                            v[j] = (v[j] & mask) | (b & ~mask);
                        }
                    }
                    u32* dw = reinterpret_cast<u32*>(dst);
                    dw[0] = v[0]; dw[1] = v[1]; dw[2] = v[2];
                    if (!left) { dw[3] = v[3]; dw[4] = v[4]; dw[5] = v[5]; }
                }
                __syncwarp();
                const int lc = wl[5 * 12 + 5];
                // SAD of the centre-subtracted patches for the 11 shifts: |(L - Lc) - (R - Rc_inc)| = |(L - Lc + Rc_inc) - R|, one
                // add and one absolute-difference-accumulate per (pixel, shift)
                int rc[11];
#pragma unroll
                for (int inc = 0; inc < 11; ++inc) rc[inc] = (int)wr[5 * 24 + inc + 5] - lc;
                int dist[11];
#pragma unroll
                for (int inc = 0; inc < 11; ++inc) dist[inc] = 0;
                for (int t = lane; t < 121; t += 32) {        // this lane's pixels; all 11 shifts per pixel
                    const int r = t / 11, cc = t - r * 11;
                    const int lv = wl[r * 12 + cc];
                    const unsigned char* rp = wr + r * 24 + cc;
#pragma unroll
                    for (int inc = 0; inc < 11; ++inc) dist[inc] = (int)__sad(lv + rc[inc], (int)rp[inc], (unsigned)dist[inc]);
                }
#pragma unroll
                for (int inc = 0; inc < 11; ++inc) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) dist[inc] += __shfl_xor_sync(0xffffffffu, dist[inc], o);
                }
                __syncwarp();
                int bestInc = 0, bestSad = dist[0];
#pragma unroll
                for (int inc = 1; inc < 11; ++inc) if (dist[inc] < bestSad) { bestSad = dist[inc]; bestInc = inc; }
                if (bestInc != 0 && bestInc != 10) {                                    // Frame.py:257
                    float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
                    for (int inc = 1; inc < 10; ++inc) if (inc == bestInc) { d1 = (float)dist[inc - 1]; d2 = (float)dist[inc]; d3 = (float)dist[inc + 1]; }
                    const float deltaR = __fdiv_rn(d1 - d3, 2.0f * (d1 + d3 - 2.0f * d2));   // Frame.py:264
                    if (!(deltaR < -1.f || deltaR > 1.f)) {
                        const float bestuR = SG.sf[oL] * ((float)(sr + bestInc - 5) + deltaR);   // Frame.py:269
                        const float disparity = uL - bestuR;
                        if (0.f <= disparity && disparity < A.maxD) {                    // Frame.py:272
                            outS = bestSad;                                              // vDistIdx.append((bestDist, iL)), :279
                            if (disparity <= 0.f) {                                      // python-float branch, :273-275
                                outD = (float)(A.mbf / 0.01);
                                outU = (float)((double)uL - 0.01);
                            } else {
                                outD = __fdiv_rn(A.mbf32, disparity);
                                outU = bestuR;
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) {
            A.uRight[oi] = outU;
            A.depth[oi] = outD;
            if (A.matchIdx) A.matchIdx[oi] = outM;
            if (A.sadDist) A.sadDist[oi] = outS;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Optional extension (NOT in the reference, whose vDistIdx is dead code -- SURVEY.md F7): upstream ORB-SLAM2's
// median-distance outlier cull at the end of ComputeStereoMatches: with the accepted matches' SAD minima sorted,
// median = element [n / 2], thDist = 1.5f * 1.4f * median, every match with dist >= thDist is dropped.
// One CTA per pair; the median is found by a two-level counting select on the integer SAD values (<= 61 710).
// ------------------------------------------------------------------------------------------------
#define MC_THREADS 256
__global__ void __launch_bounds__(MC_THREADS) k_median_cull(const int* __restrict__ nL, int n_stride, int out_stride,
                                                            const int* __restrict__ sadDist, float* __restrict__ uRight,
                                                            float* __restrict__ depth) {
    pdl_enter();
    __shared__ int hist[256];
    __shared__ int sel[3];     // [0] matches, [1] selected high byte, [2] rank inside it
    const int pair = blockIdx.x, n = nL[(size_t)pair * n_stride];
    const int* sd = sadDist + (size_t)pair * out_stride;
    for (int i = threadIdx.x; i < 256; i += MC_THREADS) hist[i] = 0;
    if (threadIdx.x == 0) sel[0] = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += MC_THREADS) {
        const int d = sd[i];
        if (d >= 0) { atomicAdd(&hist[min(d >> 8, 255)], 1); ++mine; }
    }
    atomicAdd(&sel[0], mine);
    __syncthreads();
    const int m = sel[0];
    if (m == 0) return;
    if (threadIdx.x == 0) {
        int rank = m / 2, b = 0;
        while (rank >= hist[b]) { rank -= hist[b]; ++b; }
        sel[1] = b; sel[2] = rank;
    }
    __syncthreads();
    const int hb = sel[1];
    const int rank = sel[2];
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += MC_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += MC_THREADS) {
        const int d = sd[i];
        if (d >= 0 && min(d >> 8, 255) == hb) atomicAdd(&hist[d & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r = rank, b = 0;
        while (r >= hist[b]) { r -= hist[b]; ++b; }
        sel[1] = (hb << 8) | b;
    }
    __syncthreads();
    const float thDist = (1.5f * 1.4f) * (float)sel[1];
    for (int i = threadIdx.x; i < n; i += MC_THREADS) {
        const int d = sd[i];
        if (d >= 0 && !((float)d < thDist)) { uRight[(size_t)pair * out_stride + i] = -1.f; depth[(size_t)pair * out_stride + i] = -1.f; }
    }
}

// ------------------------------------------------------------------------------------------------
// SURVEY.md 8(f) rank 2 -- BoW transform of the descriptors: the tree descent of
// pyDBoW/TemplatedVocabulary.py:139-163 (transform_feature).  One warp per descriptor; at every level lane c
// computes the Hamming distance to child c (FORB.distance, pyDBoW/FORB.py:31-33, as __popc over 2 x 16-byte loads)
// and the warp takes the minimum of (dist << 8 | c): the reference's "first strict minimum in child order".
// Outputs per feature: the leaf it ends in and the node it passes at depth `nid_level` (-1 if its path is shorter);
// the weighting / dictionary assembly (and the reference's stale-node-id quirk) stay on the host.
// ------------------------------------------------------------------------------------------------
#define VOC_WARPS 8
__global__ void __launch_bounds__(VOC_WARPS * 32) k_vocab_descend(const u8* __restrict__ desc, int n, const int* __restrict__ child_begin,
                                                                  const int* __restrict__ child_ids, const u8* __restrict__ node_desc,
                                                                  int nid_level, int* __restrict__ leaf_node, int* __restrict__ level_node) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * VOC_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4* dp = reinterpret_cast<const uint4*>(desc + (size_t)i * 32);
    const uint4 f0 = dp[0], f1 = dp[1];
    int node = 0, level = 0, at_level = -1;
    for (;;) {
        const int cb = child_begin[node], ce = child_begin[node + 1];
        if (cb == ce) break;                                  // is_leaf()
        unsigned best = 0xffffffffu;
        for (int c = cb + lane; c < ce; c += 32) {
            const int id = child_ids[c];
            const uint4* np = reinterpret_cast<const uint4*>(node_desc + (size_t)id * 32);
            const uint4 a = __ldg(np), b = __ldg(np + 1);
            const unsigned d = __popc(f0.x ^ a.x) + __popc(f0.y ^ a.y) + __popc(f0.z ^ a.z) + __popc(f0.w ^ a.w) +
                               __popc(f1.x ^ b.x) + __popc(f1.y ^ b.y) + __popc(f1.z ^ b.z) + __popc(f1.w ^ b.w);
            best = min(best, (d << 16) | (unsigned)(c - cb));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        node = child_ids[cb + (int)(best & 0xffff)];
        ++level;
        if (level == nid_level) at_level = node;
    }
    if (lane == 0) { leaf_node[i] = node; level_node[i] = at_level; }
}

// ------------------------------------------------------------------------------------------------
// SURVEY.md 8(f) rank 1 (first piece) -- all-pairs Hamming distances between two descriptor sets, the quantity
// ORBMatcher.descriptor_distance (ORBMatcher.py:12-14) computes one pair at a time in Python inside the
// search_by_BoW_* loops.  Thread j holds descriptor B_j in registers and walks HM_ROWS rows of A (warp-uniform
// loads); distances are written as uint16 (0..256), coalesced along j.
// ------------------------------------------------------------------------------------------------
#define HM_ROWS 16
__global__ void __launch_bounds__(128) k_hamming_matrix(const u8* __restrict__ A, int nA, const u8* __restrict__ B, int nB,
                                                        unsigned short* __restrict__ out) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int i0 = blockIdx.y * HM_ROWS;
    if (j >= nB) return;
    const uint4* bp = reinterpret_cast<const uint4*>(B + (size_t)j * 32);
    const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
#pragma unroll 4
    for (int i = i0; i < min(i0 + HM_ROWS, nA); ++i) {
        const uint4* ap = reinterpret_cast<const uint4*>(A + (size_t)i * 32);
        const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
        const unsigned d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
        out[(size_t)i * nB + j] = (unsigned short)d;
    }
}


// ------------------------------------------------------------------------------------------------
// Frame.get_features_in_area (Frame.py:373-416) for a batch of queries, fused with the Hamming distance
// the projection searches take to each feature it returns (ORBMatcher.py:255, 355).  One warp per query walks the query's grid
// cells in the reference's order (ix outer, iy inner, features of a cell in ascending index = the order
// assign_features_to_grid appended them, Frame.py:153-159), applies the octave window and the |dx| < r, |dy| < r test, and
// emits (feature index, distance) in that order with ballot-ranked stores.  The caller evaluates the cell range with the
// reference's own scalar arithmetic; F is the floating type that arithmetic ran in (NumPy >= 2 keeps float32 when the query
// coordinate is a float32 scalar and the other operands are Python floats, otherwise it is float64).
// count_only: pass 1 writes the per-query counts, pass 2 (after a host prefix sum) the entries.
// ------------------------------------------------------------------------------------------------
#define AQ_WARPS 8
template <typename F>
__global__ void __launch_bounds__(AQ_WARPS * 32) k_area_hamming(int M, const double* __restrict__ qxyr, const int* __restrict__ qlvl,
                                                                const int* __restrict__ qcell, const u8* __restrict__ qdesc, int rows,
                                                                const int* __restrict__ cellStart, const int* __restrict__ cellIdx,
                                                                const float* __restrict__ kxy, const int* __restrict__ koct,
                                                                const u8* __restrict__ kdesc, int count_only, int* __restrict__ qcount,
                                                                const int* __restrict__ qstart, int* __restrict__ outIdx,
                                                                int* __restrict__ outDist) {
    const int q = blockIdx.x * AQ_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= M) return;
    const F x = (F)qxyr[3 * q], y = (F)qxyr[3 * q + 1], r = (F)qxyr[3 * q + 2];
    const int minL = qlvl[2 * q], maxL = qlvl[2 * q + 1];
    const bool checkLevels = minL > 0 || maxL >= 0;
    const int c0 = qcell[4 * q], c1 = qcell[4 * q + 1], r0 = qcell[4 * q + 2], r1 = qcell[4 * q + 3];
    uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
    if (!count_only) {
        const uint4* dq = reinterpret_cast<const uint4*>(qdesc + (size_t)q * 32);
        d0 = __ldg(dq); d1 = __ldg(dq + 1);
    }
    const u32 lt = (1u << lane) - 1;
    int n = 0;
    const int base = count_only ? 0 : qstart[q];
    for (int ix = c0; ix <= c1; ++ix) {
        // the cells (ix, r0..r1) are consecutive in the CSR: one contiguous run of feature slots per grid column
        const int beg = cellStart[ix * rows + r0], end = cellStart[ix * rows + r1 + 1];
        for (int k = beg; k < end; k += 32) {
            const int i = k + lane;
            bool in = false;
            int g = 0;
            if (i < end) {
                g = cellIdx[i];
                const int o = koct[g];
                in = !(checkLevels && (o < minL || (maxL >= 0 && o > maxL)));
                const F dx = (F)kxy[2 * g] - x, dy = (F)kxy[2 * g + 1] - y;
                in = in && (dx < 0 ? -dx : dx) < r && (dy < 0 ? -dy : dy) < r;
            }
            const u32 m = __ballot_sync(0xffffffffu, in);
            if (in && !count_only) {
                const uint4* dk = reinterpret_cast<const uint4*>(kdesc + (size_t)g * 32);
                const uint4 a = __ldg(dk), b = __ldg(dk + 1);
                const int pos = base + n + __popc(m & lt);
                outIdx[pos] = g;
                outDist[pos] = __popc(a.x ^ d0.x) + __popc(a.y ^ d0.y) + __popc(a.z ^ d0.z) + __popc(a.w ^ d0.w) +
                               __popc(b.x ^ d1.x) + __popc(b.y ^ d1.y) + __popc(b.z ^ d1.z) + __popc(b.w ^ d1.w);
            }
            n += __popc(m);
        }
    }
    if (count_only && lane == 0) qcount[q] = n;
}
