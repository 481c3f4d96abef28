// Image-domain kernels: K1 pyramid (border + resize chain), K5 Gaussian blur, K2 per-cell FAST.
// All integer / byte work, HBM- and shared-memory-bound; no tensor cores by design (no dense contraction).
#pragma once
#include <cuda.h>      // CUtensorMap (the FAST windows are 2-D TMA boxes)
#include "plan.h"

typedef unsigned char u8;
typedef unsigned int u32;

// Programmatic dependent launch (every kernel of the launch sequence is launched with cudaLaunchAttributeProgrammaticStreamSerialization):
// `launch_dependents` at the top of a CTA lets the NEXT kernel of the stream start placing its CTAs as soon as every CTA of this grid
// has started, i.e. into the slots the last, partially filled wave leaves free; `wait` blocks until the previous kernel has completed
// and flushed, so nothing of its output is read early.  The launch latency and ramp-up of a kernel then overlap the tail of its
// predecessor (14 kernel boundaries per chunk).
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ int reflect101(int p, int len) {
    // OpenCV borderInterpolate(BORDER_REFLECT_101); loops only when the border is wider than the image
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do { p = p < 0 ? -p : 2 * len - 2 - p; } while ((unsigned)p >= (unsigned)len);
    return p;
}

// ------------------------------------------------------------------------------------------------
// K1a: level 0 = input image + 19-px BORDER_REFLECT_101 (ComputePyramid, ORBextractor.cpp:1125-1129).
// imgA holds slots [0, splitA), imgB the rest (left / right image batches).
// ------------------------------------------------------------------------------------------------
// One thread writes 16 consecutive bytes of the bordered buffer (one 128-bit store).  Interior spans read five
// aligned 32-bit words of the (arbitrarily aligned) source row and funnel-shift them into place; the spans that touch the
// reflected rim (two or three per row on either side) are byte work.  The two kinds never share a warp: blocks
// [0, blkA) own the interior spans [vA_lo, vA_lo + vA_n) of every row, the remaining blocks own the rim spans -- a warp
// that held one rim thread used to pay the ~300-instruction byte path for all 32 lanes (ncu: 10 active threads per
// instruction on average).
__global__ void __launch_bounds__(256) k_border0(const __grid_constant__ Plan P, const u8* __restrict__ imgA,
                                                 const u8* __restrict__ imgB, int splitA, u8* __restrict__ pyr,
                                                 int blkA, int vA_lo, int vA_n, u32 magicA, int vB_n, u32 magicB, int* __restrict__ fast_counter) {
    pdl_enter();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *fast_counter = 0;     // k_fast_cells' cell counter of this launch sequence
    const LevelGeom& G = P.lv[0];
    const int slot = blockIdx.y;
    const bool interior = (int)blockIdx.x < blkA;
    const int idx = (interior ? (int)blockIdx.x : (int)blockIdx.x - blkA) * (int)blockDim.x + (int)threadIdx.x;
    const int per_row = interior ? vA_n : vB_n;
    int by = (int)__umulhi((u32)idx, interior ? magicA : magicB);     // idx / per_row via ceil(2^32 / d); may overshoot by one
    if (by * per_row > idx) --by;
    if (by >= G.rows) return;
    int v = idx - by * per_row;
    v = interior ? vA_lo + v : (v < vA_lo ? v : v + vA_n);
    const int bx = v << 4;
    const u8* img = slot < splitA ? imgA + (size_t)slot * P.H * P.W : imgB + (size_t)(slot - splitA) * P.H * P.W;
    const u8* srow = img + (size_t)reflect101(by - ORB_EDGE, P.H) * P.W;
    const int x0 = bx - ORB_EDGE;
    u32 o[4] = {0u, 0u, 0u, 0u};
    if (interior) {                                              // x0 >= 0 && x0 + 19 < W (the 5th word may touch 3 bytes past the span)
        const size_t addr = reinterpret_cast<size_t>(srow + x0);
        const u32* wp = reinterpret_cast<const u32*>(addr & ~(size_t)3);
        const int sh = 8 * (int)(addr & 3);
        const u32 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
        o[0] = __funnelshift_r(w0, w1, sh); o[1] = __funnelshift_r(w1, w2, sh);
        o[2] = __funnelshift_r(w2, w3, sh); o[3] = __funnelshift_r(w3, w4, sh);
    } else {
        const int bw = G.w + 2 * ORB_EDGE;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int x = bx + k;
            const u32 b = x < bw ? srow[reflect101(x - ORB_EDGE, P.W)] : 0;
            o[k >> 2] |= b << (8 * (k & 3));
        }
    }
    *reinterpret_cast<uint4*>(pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs + (size_t)by * G.pitch + bx) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// K1b: level l = cv::resize(level l-1 ROI, INTER_LINEAR) + in-place reflect-101 border
// (ORBextractor.cpp:1118-1123).  Border pixels are computed directly as the resized value at the
// reflected coordinate, so one launch writes the whole bordered buffer (no second pass).
// Fixed point exactly as OpenCV's 8-bit linear path: 11-bit coefficients,
// dst = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2   (SURVEY.md App. A2).
// ------------------------------------------------------------------------------------------------
// The tables are indexed by BORDERED coordinates (the host applied the reflect-101 map and padded each row of
// entries to the buffer pitch).  One thread = 4 columns x RS_ROWS rows of the bordered output; its XGroup entry (one
// aligned 32-byte read) says where the 12-byte source window of the 4 columns starts, and each source row it needs is
// three aligned 32-bit words -> two funnel shifts (8 bytes from column 0's left sample on) -> per column one PRMT that
// picks (p[sx], p[sx+1]) and one 2-way dot product (DP2A) against the packed 11-bit coefficients.  Horizontally filtered
// source rows are kept for the next output row (at scale 1.2 consecutive output rows share one), so a thread filters
// ~1.4 source rows per output row.  Groups the host marked (reflected border columns, scale factors >= 2) take the
// per-byte path through XTab.
// Vertical step of cv::resize's fixed-point bilinear: ((b0 * (T0 >> 4) >> 16) + (b1 * (T1 >> 4) >> 16) + 2) >> 2.
// b0s / b1s are the 11-bit row weights pre-shifted left by 16, so each ">> 16" product is one multiply-high with
// accumulate (IMAD.HI, on the FMA pipe -- these kernels are bound by the ALU pipe).  The result is at most 255
// (b0 + b1 <= 2049, T >> 4 <= 32655), so no byte mask is needed.
// ------------------------------------------------------------------------------------------------
#define RS_ROWS 4
__device__ __forceinline__ u32 rs_vert(int T0, int T1, u32 b0s, u32 b1s) {
    return (__umulhi((u32)T0 >> 4, b0s) + __umulhi((u32)T1 >> 4, b1s) + 2u) >> 2;
}
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) {       // raw PRMT: the selector's upper bits are ignored by the hardware
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// Thread -> work mapping: blocks [0, blkA) own the interior column groups [gA_lo, gA_lo + gA_n) of every row group (word-window
// path, fully converged warps); the remaining blocks own the gB_n groups per row that touch the reflected border columns
// (per-byte path).  Mixed warps used to execute both paths (ncu: 16-25 active threads per instruction).
// one work item = 4 columns x RS_ROWS rows of the bordered output of level l (interior: word-window path, else per-byte path)
__device__ __forceinline__ void resize_item(const Plan& P, int l, bool interior, int rq, int gx, int slot, u8* __restrict__ pyr,
                                            const XTab* __restrict__ xtab, const XGroup* __restrict__ xgrp, const YTab* __restrict__ ytab) {
    const LevelGeom& G = P.lv[l];
    const LevelGeom& S = P.lv[l - 1];
    const int by0 = rq * RS_ROWS, bx = gx << 2;
    if (by0 >= G.rows) return;
    u8* base = pyr + (size_t)slot * P.pyr_bytes;
    const u8* srcb = base + S.pyr_ofs;                       // bordered buffer of level l-1 (ROI at +19, +19)
    const uint2* yt = reinterpret_cast<const uint2*>(ytab + G.ytab_ofs + by0);      // {y0 | y1 << 16, b0 | b1 << 16}
    const int4* gp = reinterpret_cast<const int4*>(xgrp + (G.xtab_ofs >> 2) + gx);
    const int4 g0 = __ldg(gp), g1 = __ldg(gp + 1);           // {wofs, shift8, sel01, sel23}, {cf0..cf3}
    u8* dst = base + G.pyr_ofs + (size_t)by0 * G.pitch + bx;
    const int nr = min(RS_ROWS, G.rows - by0);
    if (interior) {
        const u32 sh = (u32)g0.y, s0 = (u32)g0.z, s1 = (u32)g0.z >> 16, s2 = (u32)g0.w, s3 = (u32)g0.w >> 16;
        const u32 spw = (u32)S.pitch >> 2;
        const u32* wsrc = reinterpret_cast<const u32*>(srcb) + g0.x + ORB_EDGE * spw;
        // all loads first (the table rows, then every source row the RS_ROWS outputs need), arithmetic afterwards: the
        // kernel is latency-bound otherwise.  Row r's upper source row is usually row r-1's lower one (scale 1.2) and is
        // only fetched when it is not.
        uint2 yv[RS_ROWS];
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) yv[r] = __ldg(yt + min(r, nr - 1));
        u32 wb[RS_ROWS][3], wa[RS_ROWS][3];
        bool la[RS_ROWS];
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
            const u32* q = wsrc + (yv[r].x >> 16) * spw;           // y1: clamped source rows are never negative
            wb[r][0] = q[0]; wb[r][1] = q[1]; wb[r][2] = q[2];
        }
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
            la[r] = r == 0 || (yv[r].x & 0xffffu) != (yv[r - 1].x >> 16);
            wa[r][0] = wa[r][1] = wa[r][2] = 0u;
            if (la[r]) {
                const u32* q = wsrc + (yv[r].x & 0xffffu) * spw;   // y0
                wa[r][0] = q[0]; wa[r][1] = q[1]; wa[r][2] = q[2];
            }
        }
        auto hrow = [&](const u32 (&w)[3], int (&T)[4]) {
            const u32 lo = __funnelshift_r(w[0], w[1], sh), hi = __funnelshift_r(w[1], w[2], sh);   // 8 bytes from column 0's left sample on
            T[0] = (int)__dp2a_lo((u32)g1.x, prmt(lo, hi, s0), 0u);                                  // p[sx] * a0 + p[sx+1] * a1
            T[1] = (int)__dp2a_lo((u32)g1.y, prmt(lo, hi, s1), 0u);
            T[2] = (int)__dp2a_lo((u32)g1.z, prmt(lo, hi, s2), 0u);
            T[3] = (int)__dp2a_lo((u32)g1.w, prmt(lo, hi, s3), 0u);
        };
        int A[4], B[4] = {0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
            if (la[r]) {
                hrow(wa[r], A);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) A[j] = B[j];
            }
            hrow(wb[r], B);
            const u32 b0s = yv[r].y << 16, b1s = yv[r].y & 0xffff0000u;
            u32 v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) v |= rs_vert(A[j], B[j], b0s, b1s) << (8 * j);
            if (r < nr) *reinterpret_cast<u32*>(dst + (size_t)r * G.pitch) = v;
        }
    } else {
        const u8* roi = srcb + (size_t)ORB_EDGE * S.pitch + ORB_EDGE;
        const XTab* xt = xtab + G.xtab_ofs + bx;
        const int bw = G.w + 2 * ORB_EDGE;
        for (int r = 0; r < nr; ++r) {
            const uint2 y = __ldg(yt + r);
            const int y0 = (int)(y.x & 0xffffu), y1 = (int)(y.x >> 16);
            const u32 b0s = y.y << 16, b1s = y.y & 0xffff0000u;
            u32 v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (bx + k < bw) {
                    // when sx is the last column a1 == 0 and sx+1 reads the (valid) border pixel
                    const int sx = xt[k].sx, a0 = xt[k].a0, a1 = xt[k].a1;
                    const u8* r0 = roi + (size_t)y0 * S.pitch + sx;
                    const u8* r1 = roi + (size_t)y1 * S.pitch + sx;
                    v |= rs_vert(r0[0] * a0 + r0[1] * a1, r1[0] * a0 + r1[1] * a1, b0s, b1s) << (8 * k);
                }
            }
            *reinterpret_cast<u32*>(dst + (size_t)r * G.pitch) = v;
        }
    }
}

__global__ void __launch_bounds__(256, 5) k_resize(const __grid_constant__ Plan P, int l, int blkA, int gA_lo, int gA_n, u32 magicA,
                                                int gB_n, u32 magicB, u8* __restrict__ pyr,
                                                const XTab* __restrict__ xtab, const XGroup* __restrict__ xgrp,
                                                const YTab* __restrict__ ytab) {
    pdl_enter();
    const bool interior = (int)blockIdx.x < blkA;
    const int idx = (interior ? (int)blockIdx.x : (int)blockIdx.x - blkA) * (int)blockDim.x + (int)threadIdx.x;
    const int per_row = interior ? gA_n : gB_n;
    int rq = (int)__umulhi((u32)idx, interior ? magicA : magicB);   // row group = idx / per_row via ceil(2^32 / d); may overshoot by one
    if (rq * per_row > idx) --rq;
    int gx = idx - rq * per_row;
    gx = interior ? gA_lo + gx : (gx < gA_lo ? gx : gx + gA_n);
    resize_item(P, l, interior, rq, gx, blockIdx.y, pyr, xtab, xgrp, ytab);
}

// ------------------------------------------------------------------------------------------------
// K5: GaussianBlur 7x7 sigma 2 of every level (ORBextractor.cpp:1084-1085), OpenCV fixed-point result:
// kernel [18,34,48,56,48,34,18]/256 per axis, dst = (sum + 2^15) >> 16.
// The bordered pyramid already holds the reflect-101 halo the blur needs.
// Register-resident separable filter, no shared memory.  A warp owns a strip of 30 output words (120 columns; lane = one 32-bit
// word of the bordered buffer, lanes 30 / 31 only feed their left neighbours) and walks BLUR_RB rows down.
//   * VERTICAL pass first, on raw bytes: every four input rows are byte-transposed per lane (8 PRMT) into four registers that
//     hold one column's four rows each, so a column's seven vertical taps are two or three 4-way dot products (DP4A) against
//     constant coefficient words -- 2.5 per column and output row instead of the 4 instructions of a 16-bit formulation.
//   * HORIZONTAL pass on the 16-bit column sums, packed in pairs: the level ROI starts at buffer column 19, so output word k needs
//     the sums of words k+4 .. k+6, i.e. the lane's own two pairs, both pairs of lane + 1 and the first pair of lane + 2 (three
//     shuffles); every output is four 2-way dot products (DP2A).  Integer sums are exact in either order, the result is OpenCV's.
// ------------------------------------------------------------------------------------------------
#define BLUR_TW 120
#define BLUR_RB 32
#define BLUR_WARPS 4
#define BLUR_TH (BLUR_RB * BLUR_WARPS)
__device__ __forceinline__ u32 blur_out4(u32 m01, u32 m23, u32 r01, u32 r23, u32 q01) {
    // taps 18 34 48 56 48 34 18 over the ten sums (m01 m23 r01 r23 q01) = columns 0..9; outputs at columns 0..3
    const u32 K01 = 18u | (34u << 8), K23 = 48u | (56u << 8), K45 = 48u | (34u << 8), K6_ = 18u;
    const u32 K_0 = 18u << 8, K12 = 34u | (48u << 8), K34 = 56u | (48u << 8), K56 = 34u | (18u << 8);
    u32 a0 = __dp2a_lo(m01, K01, 32768u); a0 = __dp2a_lo(m23, K23, a0); a0 = __dp2a_lo(r01, K45, a0); a0 = __dp2a_lo(r23, K6_, a0);
    u32 a1 = __dp2a_lo(m01, K_0, 32768u); a1 = __dp2a_lo(m23, K12, a1); a1 = __dp2a_lo(r01, K34, a1); a1 = __dp2a_lo(r23, K56, a1);
    u32 a2 = __dp2a_lo(m23, K01, 32768u); a2 = __dp2a_lo(r01, K23, a2); a2 = __dp2a_lo(r23, K45, a2); a2 = __dp2a_lo(q01, K6_, a2);
    u32 a3 = __dp2a_lo(m23, K_0, 32768u); a3 = __dp2a_lo(r01, K12, a3); a3 = __dp2a_lo(r23, K34, a3); a3 = __dp2a_lo(q01, K56, a3);
    // every sum is < 2^24: the result bytes are byte 2 of each accumulator
    return __byte_perm(__byte_perm(a0, a1, 0x0062), __byte_perm(a2, a3, 0x0062), 0x5410);
}
__global__ void __launch_bounds__(BLUR_WARPS * 32) k_blur(const __grid_constant__ Plan P, const u8* __restrict__ pyr, u8* __restrict__ blur) {
    pdl_enter();
    const int slot = blockIdx.y, bid = (int)blockIdx.x;
    int l = 0;
    while (l + 1 < P.nlevels && bid >= P.lv[l + 1].blur_cta_ofs) ++l;
    const LevelGeom& G = P.lv[l];
    const int t = bid - G.blur_cta_ofs;
    const int ty = t / G.blur_tiles_x, tx = t - ty * G.blur_tiles_x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = tx * BLUR_TW + 4 * lane;                 // first of this lane's 4 output columns (ROI coords)
    const int y0 = ty * BLUR_TH + warp * BLUR_RB;           // first output row of this warp
    if (y0 >= G.h) return;                                  // warp-uniform
    const u32* src = reinterpret_cast<const u32*>(pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs);
    const int wpr = G.pitch >> 2;                            // words per buffer row
    // input bytes for outputs x0..x0+3: ROI x0-3 .. x0+6 = buffer columns x0+16 .. x0+25: this lane's word is (x0+16)/4
    const int w0 = min((x0 + 16) >> 2, wpr - 1);
    u8* dst = blur + (size_t)slot * P.blur_bytes + G.blur_ofs;
    // ask for all BLUR_RB + 8 input rows of the strip at once (one line per row): the unrolled loop below then runs on L1 hits
    for (int r = lane; r < BLUR_RB + 8; r += 32) {
        const int by = min(y0 + r - 3 + ORB_EDGE, G.rows - 1);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (size_t)by * wpr + min(((tx * BLUR_TW + 16) >> 2) + 16, wpr - 1)));
    }
    // input row i (0 .. BLUR_RB + 7) is ROI row y0 + i - 3; output j uses inputs j .. j + 6; inputs come in groups of four
    const u32 KA = 18u | (34u << 8) | (48u << 16) | (56u << 24);      // k0 k1 k2 k3
    const u32 KB = 48u | (34u << 8) | (18u << 16);                     // k4 k5 k6 0
    const u32 KC = (18u << 8) | (34u << 16) | (48u << 24);             // 0 k0 k1 k2
    const u32 KD = 56u | (48u << 8) | (34u << 16) | (18u << 24);       // k3 k4 k5 k6
    const u32 KE = (18u << 16) | (34u << 24);                          // 0 0 k0 k1
    const u32 KF = 48u | (56u << 8) | (48u << 16) | (34u << 24);       // k2 k3 k4 k5
    const u32 KG = 18u;                                                // k6 0 0 0
    const u32 KH = 18u << 24;                                          // 0 0 0 k0
    const u32 KI = 34u | (48u << 8) | (56u << 16) | (48u << 24);       // k1 k2 k3 k4
    const u32 KJ = 34u | (18u << 8);                                   // k5 k6 0 0
    u32 T[3][4];       // ring of byte-transposed groups: T[m % 3][c] = column c of inputs 4m .. 4m+3
#pragma unroll
    for (int m = 0; m < BLUR_RB / 4 + 2; ++m) {
        u32 A[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int by = min(y0 + 4 * m + k - 3 + ORB_EDGE, G.rows - 1);
            A[k] = src[(size_t)by * wpr + w0];
        }
        const u32 t0 = __byte_perm(A[0], A[1], 0x5140), t1 = __byte_perm(A[2], A[3], 0x5140);     // (a.b0 b.b0 a.b1 b.b1)
        const u32 t2 = __byte_perm(A[0], A[1], 0x7362), t3 = __byte_perm(A[2], A[3], 0x7362);     // (a.b2 b.b2 a.b3 b.b3)
        T[m % 3][0] = __byte_perm(t0, t1, 0x5410); T[m % 3][1] = __byte_perm(t0, t1, 0x7632);
        T[m % 3][2] = __byte_perm(t2, t3, 0x5410); T[m % 3][3] = __byte_perm(t2, t3, 0x7632);
        if (m >= 2) {      // outputs 4(m-2) .. 4(m-2)+3 from groups m-2, m-1, m
            const u32 (&G0)[4] = T[(m - 2) % 3];
            const u32 (&G1)[4] = T[(m - 1) % 3];
            const u32 (&G2)[4] = T[m % 3];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                u32 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (p == 0) v[c] = __dp4a(G1[c], KB, __dp4a(G0[c], KA, 0u));
                    else if (p == 1) v[c] = __dp4a(G1[c], KD, __dp4a(G0[c], KC, 0u));
                    else if (p == 2) v[c] = __dp4a(G2[c], KG, __dp4a(G1[c], KF, __dp4a(G0[c], KE, 0u)));
                    else v[c] = __dp4a(G2[c], KJ, __dp4a(G1[c], KI, __dp4a(G0[c], KH, 0u)));
                }
                const u32 m01 = __byte_perm(v[0], v[1], 0x5410), m23 = __byte_perm(v[2], v[3], 0x5410);     // 16-bit sums in pairs
                const u32 r01 = __shfl_down_sync(0xffffffffu, m01, 1), r23 = __shfl_down_sync(0xffffffffu, m23, 1);
                const u32 q01 = __shfl_down_sync(0xffffffffu, m01, 2);
                const u32 o = blur_out4(m01, m23, r01, r23, q01);
                const int y = y0 + 4 * (m - 2) + p;
                if (lane < 30 && y < G.h && x0 < G.w)   // blur pitch is a multiple of 64: the aligned 4-byte store may spill into padding only
                    *reinterpret_cast<u32*>(dst + (size_t)y * G.blur_pitch + x0) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA bulk-copy helpers (cp.async.bulk global -> shared with mbarrier completion; SASS: UBLKCP + SYNCS).
// Used to stage the FAST strip: the copy engine moves whole rows while the warps prepare their tiles, and nobody
// spends LSU instructions or registers on the staging.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 3-D tiled tensor copy (SASS: UTMALDG): one instruction moves a whole box {x .. x+BW, y .. y+BH, z} of a level into shared memory,
// rows packed BW bytes apart; coordinates are element indices -- the innermost one must be a multiple of 16 bytes (anything else
// raises an illegal-instruction fault on B200) --, out-of-range parts are zero-filled
__device__ __forceinline__ void tma_box_g2s(void* dst_smem, const CUtensorMap* tmap, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
struct alignas(64) LevelMaps { CUtensorMap m[ORB_MAX_LEVELS]; };     // one tensor map per pyramid level: {pitch, rows, slots} bytes
// bounded wait: a byte-count mistake must end in a trap, never in a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
    u32 done = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9/16 per 30-px cell with the iniThFAST -> minThFAST retry (ORBextractor.cpp:788-828) and
// cv::FAST's cell-confined strict non-max suppression.
// Persistent, barrier-free: every WARP is an independent worker that claims cells of the whole launch (all images, all levels) from
// a global counter; the cell's geometry comes from a host-built table.  A cell's raw window (detection area + 3-px rim) is staged in the warp's private
// shared-memory buffer by ONE tensor-map TMA copy (cp.async.bulk.tensor.3d of a BW x BH box at the cell's pixel coordinates,
// completion counted on the warp's own mbarrier); the copy of the NEXT cell is issued as soon as the current cell's last
// scoring pass and NMS are done, so it lands underneath the emission and the next cell's set-up and nobody ever waits at a
// block-wide barrier (the
// strip-per-CTA version lost 22 % of its warp time there; per-row bulk copies issued by 32 lanes cost ~10 instructions each).
// Per cell: quick reject of every pixel at both thresholds -> row bitmaps -> ordered work list -> exact corner strength ->
// strict NMS inside the cell's own detection window, at iniThFAST first and at minThFAST only if nothing survived ->
// survivors emitted in row-major order with ballot-ranked stores.
// Output per cell: count + packed (x | y<<12 | score<<24), x/y relative to the (16,16) detection origin.
// ------------------------------------------------------------------------------------------------
#define FAST_WARPS 4      // independent warps per CTA

// The 16 ring pixels of the Bresenham circle, clockwise from (0,+3) (OpenCV FAST pattern 16).
__device__ __forceinline__ void fast_ring(const u8* p, int SP, int (&q)[16]) {
    q[0] = p[3 * SP]; q[1] = p[3 * SP + 1]; q[2] = p[2 * SP + 2]; q[3] = p[SP + 3];
    q[4] = p[3]; q[5] = p[-SP + 3]; q[6] = p[-2 * SP + 2]; q[7] = p[-3 * SP + 1];
    q[8] = p[-3 * SP]; q[9] = p[-3 * SP - 1]; q[10] = p[-2 * SP - 2]; q[11] = p[-SP - 3];
    q[12] = p[-3]; q[13] = p[SP - 3]; q[14] = p[2 * SP - 2]; q[15] = p[3 * SP - 1];
}

// cheap reject: every arc of 9 contains one pixel of each antipodal pair, so a corner needs (0 or 8) AND (4 or 12)
// on the same side of the threshold band
__device__ __forceinline__ bool fast_quick(const u8* p, int SP, int t) {
    const int v = p[0], hi = v + t, lo = v - t;
    const int a = p[3 * SP], b = p[-3 * SP], c = p[3], d = p[-3];
    const bool br = ((a > hi) | (b > hi)) & ((c > hi) | (d > hi));
    const bool dk = ((a < lo) | (b < lo)) & ((c < lo) | (d < lo));
    return br | dk;
}

// The same reject for FOUR consecutive pixels of a row (a word-aligned group) from aligned 32-bit shared-memory words, for BOTH
// thresholds at once (iniThFAST and minThFAST, ORBextractor.cpp:808-815): the five operands (centre, x-3, x+3, y-3, y+3) are cut
// out of the words with funnel shifts, spread into 16-bit lanes (pixels 0|2 and 1|3) and tested two pixels per instruction with
// the packed min / max (VIMNMX.U16x2):
//   bright at t <=> min(max(a,b), max(c,d)) - v > t,   dark at t <=> v - max(min(a,b), min(c,d)) > t
// With A = mn + 0x7fff - v and B = v + 0x7fff - mx per lane (both inside [0x7f00, 0x80fe]: no carry, no borrow),
// (A - t) | (B - t) has bit 15 of a lane set <=> the pixel passes at t.  ri / rm: bit 15 / 31 = pixel of the low / high lane.
// The adds run as multiply-adds (x * one + y, `one` an opaque register holding 1) and the flag bits are gathered with
// multiply-high: this loop is bound by the ALU pipe (min/max, byte permutes, logic), the FMA pipe that executes IMAD idles.
// Ki / Km = 0x7fff7fff - t * 0x10001.
__device__ __forceinline__ u32 mad_lo(u32 a, u32 b, u32 c) {
    u32 d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void fast_quick2(u32 v, u32 a, u32 b, u32 c, u32 d, u32 Ki, u32 Km, u32 one, u32 mone, u32& ri, u32& rm) {
    const u32 mn = __vminu2(__vmaxu2(a, b), __vmaxu2(c, d));
    const u32 mx = __vmaxu2(__vminu2(a, b), __vminu2(c, d));
    const u32 X = mad_lo(v, mone, mn), Y = mad_lo(mx, mone, v);          // mn - v, v - mx per 16-bit lane (borrows cancel against K)
    ri = (mad_lo(X, one, Ki) | mad_lo(Y, one, Ki)) & 0x80008000u;
    rm = (mad_lo(X, one, Km) | mad_lo(Y, one, Km)) & 0x80008000u;
}
// -> min-threshold nibble in bits 0..3, ini-threshold nibble in bits 4..7 (bit k = pixel k passes); bits above 7 are scrap.
// ob = address of the first of the four pixels, 3 rows up; a multiple of 4: the groups are laid on the shared-memory word grid
// (the window's first column sits up to 3 px into its first group; those leading bits are masked off by fast_expand), so centre,
// y-3 and y+3 are single aligned words and x-3 / x+3 one constant funnel shift each.  (The kernel used to lay the groups on the
// window's own columns: eight loads, five variable shifts and four selects per group instead of five loads and two shifts.)
__device__ __forceinline__ u32 fast_quick4(const u8* ob, int SP3, int SP6, u32 Ki, u32 Km, u32 one, u32 mone) {
    const u32* wc = reinterpret_cast<const u32*>(ob + SP3);
    const u32 cl = wc[-1], C = wc[0], cr = wc[1];
    const u32 T = *reinterpret_cast<const u32*>(ob + SP6), B = *reinterpret_cast<const u32*>(ob);
    const u32 L = __funnelshift_r(cl, C, 8);           // bytes -3 .. 0
    const u32 R = __funnelshift_r(C, cr, 24);          // bytes  3 .. 6
    u32 ei, em, oi, om;
    fast_quick2(__byte_perm(C, 0, 0x4240), __byte_perm(T, 0, 0x4240), __byte_perm(B, 0, 0x4240),
                __byte_perm(R, 0, 0x4240), __byte_perm(L, 0, 0x4240), Ki, Km, one, mone, ei, em);            // pixels 0 | 2
    fast_quick2(__byte_perm(C, 0, 0x4341), __byte_perm(T, 0, 0x4341), __byte_perm(B, 0, 0x4341),
                __byte_perm(R, 0, 0x4341), __byte_perm(L, 0, 0x4341), Ki, Km, one, mone, oi, om);            // pixels 1 | 3
    // a word holds its two flags at bits 15 and 31; the high half of word * (2^(17+t) + 2^(3+t)) has them at bits t and t + 2
    // (plus a stray copy at bit 16 + t): t = 0 / 1 for the min flags of pixels 0|2 / 1|3, t = 4 / 5 for the ini flags
    u32 r = __umulhi(em, (1u << 17) + (1u << 3));
    r += __umulhi(om, (1u << 18) + (1u << 4));
    r += __umulhi(ei, (1u << 21) + (1u << 7));
    r += __umulhi(oi, (1u << 22) + (1u << 8));
    return r;        // the caller stores the low byte
}

// Phase 1 of a cell.  sa = the word-aligned address at or below the window's first detection pixel, cwa = columns from there to
// the window's right edge (<= 64).  A lane tests 4 consecutive pixels of a row and drops the two nibbles (one 16-bit store) into
// a per-row byte table rowq[row][quad] (8 quads per row for up to 32 columns, 16 beyond).  The Q = ceil(cwa / 4) quads of all rows are
// dealt to the lanes as one sequence (item = row * Q + quad, lane + 32 * step): no lane idles whatever Q is.
__device__ __forceinline__ void fast_phase1_bits(const u8* sa, int SP, int cwa, int ch, int Q, u32 rcp, int dr, int Ti, int Tm, u32 one,
                                                 u8* rowq, int lane) {
    const int sh = cwa > 32 ? 4 : 3;
    const u32 ki = 0x7fff7fffu - (u32)Ti * 0x00010001u, km = 0x7fff7fffu - (u32)Tm * 0x00010001u, mone = 0u - one;
    // Q = ceil(cwa / 4) in 1 .. 16, rcp = ceil(65536 / Q): n / Q == (n * rcp) >> 16 for n <= 32, dr = 32 / Q (all from the cell table)
    const int row = (int)(((u32)lane * rcp) >> 16), dq = 32 - dr * Q, SP3 = 3 * SP, SP6 = 6 * SP;
    int quad = lane - row * Q;
    const u8* ob = sa - SP3 + row * SP + 4 * quad;         // the lane's group, 3 rows up (the y-3 operand)
    const u8* const oend = sa - SP3 + ch * SP;
    u8* rb = rowq + (row << sh) + quad;
    const int stepO = dr * SP + 4 * dq, stepR = (dr << sh) + dq, stepOw = stepO + SP - 4 * Q, stepRw = stepR + (1 << sh) - Q;
#pragma unroll 2
    while (ob < oend) {
        *rb = (u8)fast_quick4(ob, SP3, SP6, ki, km, one, mone);
        quad += dq;
        const bool wrap = quad >= Q;
        ob += wrap ? stepOw : stepO; rb += wrap ? stepRw : stepR; quad -= wrap ? Q : 0;
    }
    __syncwarp();
}

// Row-major work list of a cell from the phase-1 table: lane r turns the masks of rows r and r + 32 into list entries (y << 6 | x)
// at the offsets an exclusive warp scan of the row counts gives -- row-major order by construction, no per-pixel ballot.
// x counts from the aligned origin (see fast_phase1_bits): colmask keeps columns [a, cwa).  Only rows [r0, r1) are listed.
// useMin = false: pixels passing at iniThFAST;  true: pixels passing at minThFAST.
// Returns the number of entries; when that exceeds the list capacity LC nothing is written (the caller then walks row bands).
__device__ __forceinline__ int fast_expand(const u8* rowq, int cwa, int ch, unsigned long long colmask, bool useMin, int r0, int r1,
                                           unsigned short* list, int LC, int lane) {
    const bool wide = cwa > 32;
    const int nsh = useMin ? 0 : 4;
    // a byte holds one quad: min nibble | ini nibble << 4; squeeze the chosen nibbles of a word's four quads into 16 bits
    auto nib4 = [nsh](u32 w) { w = (w >> nsh) & 0x0f0f0f0fu; w = (w | (w >> 4)) & 0x00ff00ffu; return (w | (w >> 8)) & 0xffffu; };
    unsigned long long m[2] = {0ull, 0ull};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = lane + 32 * h;
        if (row < min(ch, r1) && row >= r0) {
            unsigned long long mm;
            if (wide) {
                const uint4 w = *reinterpret_cast<const uint4*>(rowq + row * 16);
                mm = (unsigned long long)(nib4(w.x) | (nib4(w.y) << 16)) | ((unsigned long long)(nib4(w.z) | (nib4(w.w) << 16)) << 32);
            } else {
                const uint2 w = *reinterpret_cast<const uint2*>(rowq + row * 8);
                mm = nib4(w.x) | (nib4(w.y) << 16);
            }
            m[h] = mm & colmask;
        }
    }
    const int c0 = __popcll(m[0]), c1 = __popcll(m[1]);
    // one inclusive scan for both halves: rows 0..31 in the low 16 bits, rows 32..63 in the high ones (a row has at most 64 entries,
    // 32 rows at most 2048: no carry between the halves)
    int incl = c0 | (c1 << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    const int n0 = tot & 0xffff, n = n0 + (tot >> 16);
    if (n > LC) return n;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && ch <= 32) break;
        unsigned short* w = list + (h ? n0 + (incl >> 16) - c1 : (incl & 0xffff) - c0);
        const int ybits = (lane + 32 * h) << 6;
        u32 lo = (u32)m[h], hi = (u32)(m[h] >> 32);
        while (lo) { *w++ = (unsigned short)(ybits | (__ffs(lo) - 1)); lo &= lo - 1; }
        while (hi) { *w++ = (unsigned short)(ybits | (__ffs(hi) + 31)); hi &= hi - 1; }
    }
    __syncwarp();
    return n;
}

// 9-of-16 segment test (strictly brighter than v+t or strictly darker than v-t on 9 contiguous ring pixels).
// The 32 comparisons are done two ring pixels at a time in 16-bit lanes (SWAR): with pair = p_k | p_{k+8} << 16,
//   pair + (0x7fff - hi) per lane sets bit 15 of a lane  <=>  p > hi        (no carry: p + 0x7fff - hi < 0x10000)
//   (0x7fff + lo) - pair per lane sets bit 15 of a lane  <=>  p < lo        (no borrow: lane stays in [0x7e01, 0x80fe])
// and one shift + one and-or drops bit 15 / bit 31 at mask positions k / k + 16.
__device__ __forceinline__ bool fast_is_corner(const u8* p, int SP, int t) {
    int q[16];
    fast_ring(p, SP, q);
    const int v = p[0];
    const u32 kb = (u32)(0x7fff - (v + t)) * 0x00010001u;
    const u32 kd = (u32)(0x7fff + (v - t)) * 0x00010001u;
    u32 mb = 0, md = 0;   // bit k = ring position k (k < 8), bit k + 16 = ring position k + 8
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 pair = (u32)q[k] | ((u32)q[k + 8] << 16);
        mb |= ((pair + kb) >> (15 - k)) & (0x00010001u << k);
        md |= ((kd - pair) >> (15 - k)) & (0x00010001u << k);
    }
    // unscramble to the circular order and double it: bits 0..15 = ring, bits 16..31 = ring again
    mb = (mb & 0xffu) | ((mb >> 8) & 0xff00u); mb |= mb << 16;
    md = (md & 0xffu) | ((md >> 8) & 0xff00u); md |= md << 16;
    u32 rb = mb & (mb >> 1); rb &= rb >> 2; rb &= rb >> 4; rb &= mb >> 8;
    u32 rd = md & (md >> 1); rd &= rd >> 2; rd &= rd >> 4; rd &= md >> 8;
    return ((rb | rd) & 0xffffu) != 0;
}

// exact OpenCV corner score (cornerScore<16>) of a pixel known to be a corner: best - 1 with
// best = max over the 16 arcs of 9 of min(v - p_k) and of min(p_k - v).
// Works on the biased differences D_k = 255 + v - p_k in [0, 510]: the bright side min(p_k - v) equals
// 255 - max(D), so no negated operand ever feeds a max -- ptxas 12.9 (sm_100a, -O1 and above) drops the
// negation when it fuses max(x, -y) chains into VIMNMX3 (caught by tests/cuda_unit/fast_unit.cu).
// Two ring positions (k, k + 8) share one register as 16-bit lanes, so position k + 8 is the lane-swapped
// register of position k, and the sliding window of 9 = 3 x 3 is two passes of the native 3-input packed
// min / max (VIMNMX3.S16x2): 16 + 16 instructions give all 16 arc minima and maxima.
__device__ __forceinline__ u32 swap16(u32 x) { return __byte_perm(x, 0, 0x1032); }
__device__ __forceinline__ int fast_corner_score(const u8* p, int SP) {
    int q[16];
    fast_ring(p, SP, q);
    const u32 bias = (u32)(255 + p[0]) * 0x00010001u;
    u32 D[10];
#pragma unroll
    for (int k = 0; k < 8; ++k) D[k] = bias - ((u32)q[k] | ((u32)q[k + 8] << 16));   // no borrow: lanes stay in [0, 510]
    D[8] = swap16(D[0]); D[9] = swap16(D[1]);
    u32 n3[14], x3[14];
#pragma unroll
    for (int k = 0; k < 8; ++k) { n3[k] = __vimin3_s16x2(D[k], D[k + 1], D[k + 2]); x3[k] = __vimax3_s16x2(D[k], D[k + 1], D[k + 2]); }
#pragma unroll
    for (int k = 8; k < 14; ++k) { n3[k] = swap16(n3[k - 8]); x3[k] = swap16(x3[k - 8]); }
    u32 n9[8], x9[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { n9[k] = __vimin3_s16x2(n3[k], n3[k + 3], n3[k + 6]); x9[k] = __vimax3_s16x2(x3[k], x3[k + 3], x3[k + 6]); }
    // best dark = max of the 16 arc minima; worst bright = min of the 16 arc maxima
    u32 bd = __vimax3_s16x2(__vimax3_s16x2(n9[0], n9[1], n9[2]), __vimax3_s16x2(n9[3], n9[4], n9[5]), __vimax3_s16x2(n9[6], n9[7], n9[7]));
    u32 wb = __vimin3_s16x2(__vimin3_s16x2(x9[0], x9[1], x9[2]), __vimin3_s16x2(x9[3], x9[4], x9[5]), __vimin3_s16x2(x9[6], x9[7], x9[7]));
    const int bestDark = max((int)(bd & 0xffff), (int)(bd >> 16));          // 255 + max_arcs min(v - p)
    const int worstBright = min((int)(wb & 0xffff), (int)(wb >> 16));       // 255 - max_arcs min(p - v)
    return max(bestDark - 255, 255 - worstBright) - 1;
}

// segment test + score in one call (0 for non-corners at t); used by the unit harness
__device__ __forceinline__ int fast_score(const u8* p, int SP, int t) {
    if (!fast_quick(p, SP, t) || !fast_is_corner(p, SP, t)) return 0;
    return fast_corner_score(p, SP);
}

// Work list entry: y << 6 | x inside the cell's detection window (both < 64), bit 15 = survives NMS.
// Phases per cell (one warp), each on FULL warps thanks to in-place ordered compaction of the list:
//   1   quick reject of all window pixels at BOTH thresholds, once                        -> per-row bit table
//   2+3 exact corner strength (packed 3-input min/max) of the pixels passing at iniThFAST -> score tile (every score >= 1);
//       entries with score >= iniThFAST stay in the list (in place)
//   4   strict 8-neighbour NMS on that list
//   R   only if nothing survived (ORBextractor.cpp:811): the same two steps on the pixels passing at minThFAST, kept at
//       score >= minThFAST (the few that had passed at iniThFAST are scored a second time: same values, no second list)
//   5   ordered emission of the survivors
// Survivors at T are exactly the strict local maxima among the pixels with score >= T: a neighbour with a lower score never
// suppresses, so scores below T left in the tile are harmless and nothing is cleared between the two attempts.
struct FastCell {            // geometry of one cell (warp-uniform)
    int slot, level, rem;    // rem = cell index inside the slot (all levels)
    int iniX, iniY, cw, ch;  // window origin (incl. the 3-px rim, level coordinates) and detection size; cw <= 0: the reference skips the cell
    int cand_ofs;            // u32 index of the cell's candidate storage inside the slot's blob
    u32 deal;                // phase-1 lane dealing: rcp | dr << 17 | Q << 23 (see fast_phase1_bits)
};
// One table entry per cell of an image (built by the host from ORBextractor.cpp:783-806, see HostPlan::celltab):
//   x = iniX | iniY << 16,  y = cw | ch << 8 | level << 16 (cw = 0: skipped cell),  z = candidate offset,  w = lane dealing of phase 1
__device__ __forceinline__ void fast_cell_geom(const uint4* __restrict__ celltab, int c, int cells_per_slot, u32 slot_magic, FastCell& g) {
    int slot = (int)__umulhi((u32)c, slot_magic);             // c / cells_per_slot via ceil(2^32 / d); may overshoot by one
    if (slot * cells_per_slot > c) --slot;
    const int rem = c - slot * cells_per_slot;
    const uint4 t = __ldg(celltab + rem);
    g.slot = slot; g.rem = rem; g.level = (int)(t.y >> 16);
    g.iniX = (int)(t.x & 0xffffu); g.iniY = (int)(t.x >> 16);
    g.cw = (int)(t.y & 0xffu); g.ch = (int)((t.y >> 8) & 0xffu);
    g.cand_ofs = (int)t.z; g.deal = t.w;
}

__global__ void __launch_bounds__(FAST_WARPS * 32, 8) k_fast_cells(const __grid_constant__ Plan P, const __grid_constant__ LevelMaps M,
                                                                   const uint4* __restrict__ celltab,
                                                                   u32* __restrict__ cand, int* __restrict__ cellcnt,
                                                                   int total_cells, int cells_per_slot, u32 slot_magic, int* __restrict__ next_cell,
                                                                   u32 one /* = 1, opaque to the compiler: see fast_quick2 */,
                                                                   int SP /*window pitch = box width*/, int SR /*window rows = box height*/,
                                                                   int TP /*tile pitch*/, int TR /*tile rows*/,
                                                                   int LC /*list capacity*/, int RQ /*bytes of the row table*/, int WS /*bytes per warp*/) {
    pdl_enter();
    extern __shared__ __align__(128) u8 smem[];
    __shared__ unsigned long long win_bar[FAST_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1;
    // the warp's private carve-out (128-byte aligned): window | score tile | work list | row table
    u8* win = smem + (size_t)warp * WS;
    u8* tile = win + SP * SR;
    unsigned short* list = reinterpret_cast<unsigned short*>(tile + TP * TR);
    u8* rowq = reinterpret_cast<u8*>(list + LC);
    unsigned long long* bar = &win_bar[warp];
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    // dynamic work distribution: a warp takes the next unclaimed cell (global counter, zeroed by this sequence's k_border0), so the
    // launch ends when the cells run out, not when the unluckiest static share does
    auto claim = [&]() { int v = 0; if (lane == 0) v = atomicAdd(next_cell, 1); return __shfl_sync(0xffffffffu, v, 0); };
    int c = claim();
    FastCell cur;
    // stage a cell's window: one elected lane, one TMA box copy
    const CUtensorMap* const maps = M.m;          // address inside the kernel parameter block (__grid_constant__): valid for the TMA unit
    const u32 box_bytes = (u32)(SP * SR);
    auto stage = [&](const FastCell& g) {
        if (g.cw <= 0 || lane != 0) return;
        mbar_expect_tx(bar, box_bytes);
        tma_box_g2s(win, maps + g.level, (g.iniX + ORB_EDGE) & ~15, g.iniY + ORB_EDGE, g.slot, bar);   // the box must start 16-byte aligned
    };
    if (c < total_cells) { fast_cell_geom(celltab, c, cells_per_slot, slot_magic, cur); stage(cur); }
    const int iniTh = max(0, min(P.iniTh, 255)), minTh = max(0, min(P.minTh, 255));
    u32 parity = 0;
#pragma unroll 1
    while (c < total_cells) {
    int* cnt_out = cellcnt + (size_t)cur.slot * P.ncells + cur.rem;
    const int cw = cur.cw, ch = cur.ch;
    FastCell nxt;
    const int cn = claim();
    const bool more = cn < total_cells;
    if (more) fast_cell_geom(celltab, cn, cells_per_slot, slot_magic, nxt);
    if (cw <= 0) {                                          // cell the reference skips: nothing was staged for it
        if (lane == 0) *cnt_out = 0;
        if (more) stage(nxt);
        cur = nxt; c = cn;
        continue;
    }
    // zero the score tile (the rim is what "outside the window scores 0" means), then make sure the window has landed
    for (int i = lane; i < (TP * TR) >> 4; i += 32) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);   // TP, TR multiples of 4
    mbar_wait(bar, parity);
    parity ^= 1u;
    __syncwarp();
    // first detection pixel = box column xo (the box starts at the 16-byte boundary below the window); everything below counts
    // columns from the word boundary at or below it: s0 = that address, the window's columns are [xa, cwa)
    const int xo = ((cur.iniX + ORB_EDGE) & 15) + 3, xa = xo & 3, cwa = cw + xa;
    const u8* s0 = win + 3 * SP + (xo - xa);
    const unsigned long long colmask = (~0ull >> (64 - cw)) << xa;          // 1 <= cw, cwa <= 64

    // ---- phase 1, both thresholds ----
    fast_phase1_bits(s0, SP, cwa, ch, (int)(cur.deal >> 23), cur.deal & 0x1ffffu, (int)((cur.deal >> 17) & 63u), iniTh, minTh, one, rowq, lane);     // four pixels per lane
    // exact corner strength of list[0, n): every score >= 1 goes to the tile (a score of 0 can never win the strict NMS, so it is
    // dropped like a non-corner); entries with score >= tKeep are kept, compacted in place.  corner at T <=> best > T <=> score >= T
    auto score_list = [&](int n, int tKeep) {
        int kept = 0;
        for (int b = 0; b < n; b += 32) {
            const int i = b + lane;
            const int e = i < n ? list[i] : 0;
            const int y = e >> 6, x = e & 63;
            const int sc = i < n ? fast_corner_score(s0 + y * SP + x, SP) : 0;
            const bool c = sc >= tKeep;
            const u32 m = __ballot_sync(0xffffffffu, c);
            __syncwarp();
            if (sc > 0) tile[(y + 1) * TP + x + 1] = (u8)sc;
            if (c) list[kept + __popc(m & lt)] = (unsigned short)e;
            kept += __popc(m);
        }
        __syncwarp();
        return kept;
    };
    // strict 8-neighbour NMS inside the window: survivors get bit 15
    auto nms_list = [&](int n) {
        bool any = false;
        for (int i = lane; i < n; i += 32) {
            const int e = list[i];
            const u8* t = tile + ((e >> 6) + 1) * TP + (e & 63) + 1;
            const int sc = t[0];
            const bool keep = sc > t[-1] && sc > t[1] && sc > t[-TP - 1] && sc > t[-TP] && sc > t[-TP + 1] &&
                              sc > t[TP - 1] && sc > t[TP] && sc > t[TP + 1];
            if (keep) { list[i] = (unsigned short)(e | 0x8000); any = true; }
        }
        __syncwarp();
        return __any_sync(0xffffffffu, any);
    };
    // keep the entries whose tile score is >= tKeep (in place)
    auto filter_list = [&](int n, int tKeep) {
        int kept = 0;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            const int e = i < n ? list[i] : 0;
            const bool cc = i < n && tile[((e >> 6) + 1) * TP + (e & 63) + 1] >= tKeep;
            const u32 m = __ballot_sync(0xffffffffu, cc);
            __syncwarp();
            if (cc) list[kept + __popc(m & lt)] = (unsigned short)e;
            kept += __popc(m);
        }
        __syncwarp();
        return kept;
    };
    // ---- phase 5: ordered emission of the flagged entries of list[0, n), appended at out[count...] ----
    u32* out = cand + (size_t)cur.slot * P.cand_entries + cur.cand_ofs;
    const int xrel0 = cur.iniX - ORB_DET_ORIGIN + 3 - xa, yrel0 = cur.iniY - ORB_DET_ORIGIN + 3;
    auto emit_list = [&](int n, int count) {
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            const int e = i < n ? list[i] : 0;
            const int y = (e >> 6) & 63, x = e & 63;
            const int sc = tile[(y + 1) * TP + x + 1];
            const bool keep = (e & 0x8000) != 0;
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[count + __popc(m & lt)] = (u32)(xrel0 + x) | ((u32)(yrel0 + y) << 12) | ((u32)sc << 24);
            count += __popc(m);
        }
        return count;
    };
    // One loop body for the two attempts (a single copy of the expansion / scoring / NMS code: the kernel is sensitive to
    // instruction-cache misses):
    //   pass 0  pixels passing at iniThFAST: score, keep score >= iniThFAST, NMS (ORBextractor.cpp:808-809); done if anything survives
    //   pass 1  vKeysCell.empty() (:811): pixels passing at minThFAST: score, keep score >= minThFAST, NMS (:813-815)
    // The list holds LC entries (half the window).  A cell with more passing pixels (noise, dense texture at a low threshold) is
    // walked in bands of LC / 64 rows: first every band is scored into the tile, then every band is listed again, cut at the
    // threshold by its tile scores, suppressed and emitted -- bands in row order, so the output order is the same.
    int nB = 0, count = -1;         // count >= 0: the banded path has emitted already
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int tKeep = max(pass == 0 ? iniTh : minTh, 1);
        const int n = fast_expand(rowq, cwa, ch, colmask, pass != 0, 0, 64, list, LC, lane);
        if (n <= LC) {
            nB = score_list(n, tKeep);
            if (nms_list(nB) || minTh >= iniTh) break;
        } else {
            const int rpb = LC >> 6;
#pragma unroll 1
            for (int r0 = 0; r0 < ch; r0 += rpb) score_list(fast_expand(rowq, cwa, ch, colmask, pass != 0, r0, r0 + rpb, list, LC, lane), 256);
            count = 0;
#pragma unroll 1
            for (int r0 = 0; r0 < ch; r0 += rpb) {
                const int nb = filter_list(fast_expand(rowq, cwa, ch, colmask, pass != 0, r0, r0 + rpb, list, LC, lane), tKeep);
                nms_list(nb);
                count = emit_list(nb, count);
                __syncwarp();
            }
            if (count > 0 || minTh >= iniTh) break;
            count = -1;
        }
    }
    __syncwarp();
    if (more) stage(nxt);        // nothing reads the window any more: the next cell's copy may start now
    if (count < 0) count = emit_list(nB, 0);
    if (lane == 0) *cnt_out = count;
    __syncwarp();
    cur = nxt; c = cn;
    }
}
