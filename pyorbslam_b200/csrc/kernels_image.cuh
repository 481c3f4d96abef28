// Image-domain kernels: K1 pyramid (border + resize chain), K5 Gaussian blur, K2 per-cell FAST.
// All integer / byte work, HBM- and shared-memory-bound; no tensor cores by design (no dense contraction).
#pragma once
#include "plan.h"

typedef unsigned char u8;
typedef unsigned int u32;

__device__ __forceinline__ int reflect101(int p, int len) {
    // OpenCV borderInterpolate(BORDER_REFLECT_101); loops only when the border is wider than the image
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do { p = p < 0 ? -p : 2 * len - 2 - p; } while ((unsigned)p >= (unsigned)len);
    return p;
}

// ------------------------------------------------------------------------------------------------
// K1a: level 0 = input image + 19-px BORDER_REFLECT_101 (ComputePyramid, ORBextractor.cpp:1125-1129).
// One thread writes 4 consecutive bytes of the bordered buffer (one 32-bit store).
// imgA holds slots [0, splitA), imgB the rest (left / right image batches).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_border0(const __grid_constant__ Plan P, const u8* __restrict__ imgA,
                                                 const u8* __restrict__ imgB, int splitA, u8* __restrict__ pyr) {
    const LevelGeom& G = P.lv[0];
    const int slot = blockIdx.y;
    const int words_per_row = G.pitch >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= words_per_row * G.rows) return;
    const int by = idx / words_per_row, bx = (idx - by * words_per_row) << 2;
    const u8* img = slot < splitA ? imgA + (size_t)slot * P.H * P.W : imgB + (size_t)(slot - splitA) * P.H * P.W;
    const u8* srow = img + (size_t)reflect101(by - ORB_EDGE, P.H) * P.W;
    u32 v = 0;
    const int x0 = bx - ORB_EDGE;
    if (x0 >= 0 && x0 + 3 < P.W) {
        v = srow[x0] | (srow[x0 + 1] << 8) | (srow[x0 + 2] << 16) | ((u32)srow[x0 + 3] << 24);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x = bx + k;
            u32 b = x < G.w + 2 * ORB_EDGE ? srow[reflect101(x - ORB_EDGE, P.W)] : 0;
            v |= b << (8 * k);
        }
    }
    *reinterpret_cast<u32*>(pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs + (size_t)by * G.pitch + bx) = v;
}

// ------------------------------------------------------------------------------------------------
// K1b: level l = cv::resize(level l-1 ROI, INTER_LINEAR) + in-place reflect-101 border
// (ORBextractor.cpp:1118-1123).  Border pixels are computed directly as the resized value at the
// reflected coordinate, so one launch writes the whole bordered buffer (no second pass).
// Fixed point exactly as OpenCV's 8-bit linear path: 11-bit coefficients,
// dst = (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2   (SURVEY.md App. A2).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ Plan P, int l, u8* __restrict__ pyr,
                                                const XTab* __restrict__ xtab, const YTab* __restrict__ ytab) {
    const LevelGeom& G = P.lv[l];
    const LevelGeom& S = P.lv[l - 1];
    const int slot = blockIdx.y;
    const int words_per_row = G.pitch >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= words_per_row * G.rows) return;
    const int by = idx / words_per_row, bx = (idx - by * words_per_row) << 2;
    u8* base = pyr + (size_t)slot * P.pyr_bytes;
    const u8* src = base + S.pyr_ofs + (size_t)ORB_EDGE * S.pitch + ORB_EDGE;   // ROI origin of level l-1
    const YTab yt = ytab[G.ytab_ofs + reflect101(by - ORB_EDGE, G.h)];
    const u8* r0 = src + (size_t)yt.y0 * S.pitch;
    const u8* r1 = src + (size_t)yt.y1 * S.pitch;
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = bx + k;
        if (x < G.w + 2 * ORB_EDGE) {
            const XTab xt = xtab[G.xtab_ofs + reflect101(x - ORB_EDGE, G.w)];
            // when sx is the last column a1 == 0 and sx+1 reads the (valid) border pixel
            const int T0 = r0[xt.sx] * xt.a0 + r0[xt.sx + 1] * xt.a1;
            const int T1 = r1[xt.sx] * xt.a0 + r1[xt.sx + 1] * xt.a1;
            const int d = (((yt.b0 * (T0 >> 4)) >> 16) + ((yt.b1 * (T1 >> 4)) >> 16) + 2) >> 2;
            v |= (u32)(d & 0xff) << (8 * k);
        }
    }
    *reinterpret_cast<u32*>(base + G.pyr_ofs + (size_t)by * G.pitch + bx) = v;
}

// ------------------------------------------------------------------------------------------------
// K5: GaussianBlur 7x7 sigma 2 of every level (ORBextractor.cpp:1084-1085), OpenCV fixed-point result:
// kernel [18,34,48,56,48,34,18]/256 per axis, dst = (sum + 2^15) >> 16.
// The bordered pyramid already holds the reflect-101 halo the blur needs.  Tile 64 x 16 per CTA.
// ------------------------------------------------------------------------------------------------
#define BLUR_TW 64
#define BLUR_TH 16
__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ Plan P, const u8* __restrict__ pyr, u8* __restrict__ blur) {
    __shared__ u8 raw[BLUR_TH + 6][BLUR_TW + 8];
    __shared__ unsigned short hs[BLUR_TH + 6][BLUR_TW];
    const int slot = blockIdx.y;
    int l = 0;
    while (l + 1 < P.nlevels && (int)blockIdx.x >= P.lv[l + 1].blur_cta_ofs) ++l;
    const LevelGeom& G = P.lv[l];
    const int t = blockIdx.x - G.blur_cta_ofs;
    const int ty = t / G.blur_tiles_x, tx = t - ty * G.blur_tiles_x;
    const int x0 = tx * BLUR_TW, y0 = ty * BLUR_TH;
    const u8* src = pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs;
    // raw tile: ROI (x0-3 .. x0+66, y0-3 .. y0+18) -> bordered coords (+19); clamp to the buffer for partial tiles
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * (BLUR_TW + 6); i += 256) {
        const int r = i / (BLUR_TW + 6), c = i - r * (BLUR_TW + 6);
        const int by = min(y0 + r + ORB_EDGE - 3, G.rows - 1), bx = min(x0 + c + ORB_EDGE - 3, G.pitch - 1);
        raw[r][c] = src[(size_t)by * G.pitch + bx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        const u8* p = &raw[r][c];
        hs[r][c] = (unsigned short)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    // each thread: 4 consecutive x of one row -> one 32-bit store
    const int r = threadIdx.x / (BLUR_TW / 4), c4 = (threadIdx.x - r * (BLUR_TW / 4)) * 4;
    const int y = y0 + r, x = x0 + c4;
    if (y < G.h && x < G.w) {
        u32 v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c4 + k;
            const u32 acc = 18u * (hs[r][c] + hs[r + 6][c]) + 34u * (hs[r + 1][c] + hs[r + 5][c]) +
                            48u * (hs[r + 2][c] + hs[r + 4][c]) + 56u * hs[r + 3][c];
            v |= ((acc + 32768u) >> 16) << (8 * k);
        }
        // blur pitch is a multiple of 64, so the 4-byte store is aligned; bytes past w are padding
        *reinterpret_cast<u32*>(blur + (size_t)slot * P.blur_bytes + G.blur_ofs + (size_t)y * G.blur_pitch + x) = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9/16 per 30-px cell with the iniThFAST -> minThFAST retry (ORBextractor.cpp:788-828) and
// cv::FAST's cell-confined strict non-max suppression.
// One CTA = one strip of FAST_WARPS consecutive cells of a cell row: the raw strip (cells + 3-px rim) is
// staged in shared memory once with 32-bit loads, then one warp per cell computes the threshold-free
// corner score, suppresses non-maxima inside its own detection window, decides the threshold
// (post-NMS list empty at iniThFAST -> use minThFAST) and emits the survivors in row-major order with
// ballot-ranked stores.  Output per cell: count + packed (x | y<<12 | score<<24), x/y relative to (16,16).
// ------------------------------------------------------------------------------------------------
#define FAST_WARPS 8

// 9-of-16 segment test + exact OpenCV corner score (cornerScore<16>): score = best - 1 with
// best = max over the 16 arcs of 9 of min(v - p_k) and of min(p_k - v).  Returns 0 for non-corners at t.
__device__ __forceinline__ int fast_score(const u8* p, int SP, int t) {
    const int v = p[0], hi = v + t, lo = v - t;
    int q[16];
    q[0] = p[3 * SP]; q[8] = p[-3 * SP];
    int br = (q[0] > hi) | (q[8] > hi), dk = (q[0] < lo) | (q[8] < lo);
    if (!(br | dk)) return 0;              // every 9-arc contains one pixel of each antipodal pair
    q[4] = p[3]; q[12] = p[-3];
    br &= (q[4] > hi) | (q[12] > hi); dk &= (q[4] < lo) | (q[12] < lo);
    if (!(br | dk)) return 0;
    q[1] = p[3 * SP + 1]; q[2] = p[2 * SP + 2]; q[3] = p[SP + 3];
    q[5] = p[-SP + 3]; q[6] = p[-2 * SP + 2]; q[7] = p[-3 * SP + 1];
    q[9] = p[-3 * SP - 1]; q[10] = p[-2 * SP - 2]; q[11] = p[-SP - 3];
    q[13] = p[SP - 3]; q[14] = p[2 * SP - 2]; q[15] = p[3 * SP - 1];
    u32 mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (u32)(q[k] > hi) << k; md |= (u32)(q[k] < lo) << k; }
    mb |= mb << 16; md |= md << 16;
    u32 rb = mb & (mb >> 1); rb &= rb >> 2; rb &= rb >> 4; rb &= mb >> 8;
    u32 rd = md & (md >> 1); rd &= rd >> 2; rd &= rd >> 4; rd &= md >> 8;
    if (!((rb | rd) & 0xffffu)) return 0;
    // sliding min / max over windows of 9 on the circle (doubling: 2, 4, 8, then +1).
    // Works on the biased differences D = 255 + v - p_k in [0, 510]: the bright side min(p_k - v) equals
    // 255 - max(D), so no negated operand ever feeds a max -- ptxas 12.9 (sm_100a, -O1 and above) drops the
    // negation when it fuses max(x, -y) chains into VIMNMX3 (caught by tests/cuda_unit/fast_unit.cu).
    int d[16], a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = 255 + v - q[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) { a[k] = min(d[k], d[(k + 1) & 15]); b[k] = max(d[k], d[(k + 1) & 15]); }
    int a4[16], b4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { a4[k] = min(a[k], a[(k + 2) & 15]); b4[k] = max(b[k], b[(k + 2) & 15]); }
    int bestDark = 0, worstBright = 510;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int mn = min(min(a4[k], a4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx = max(max(b4[k], b4[(k + 4) & 15]), d[(k + 8) & 15]);
        bestDark = max(bestDark, mn);          // 255 + max_arcs min(v - p)
        worstBright = min(worstBright, mx);    // 255 - max_arcs min(p - v)
    }
    const int best = max(bestDark - 255, 255 - worstBright);
    return best - 1;
}

__global__ void __launch_bounds__(FAST_WARPS * 32) k_fast_cells(const __grid_constant__ Plan P, const u8* __restrict__ pyr,
                                                                u32* __restrict__ cand, int* __restrict__ cellcnt,
                                                                int SP /*strip pitch*/, int SR /*strip rows*/, int TP /*tile pitch*/, int TR /*tile rows*/) {
    extern __shared__ __align__(16) u8 smem[];
    u8* strip = smem;
    const int slot = blockIdx.y;
    int l = 0;
    while (l + 1 < P.nlevels && (int)blockIdx.x >= P.lv[l + 1].fast_cta_ofs) ++l;
    const LevelGeom& G = P.lv[l];
    const int local = blockIdx.x - G.fast_cta_ofs;
    const int ci = local / G.fast_groups, g = local - ci * G.fast_groups;
    const int j0 = g * FAST_WARPS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tLow = max(0, min(min(P.iniTh, P.minTh), 255));

    const int iniY = ORB_DET_ORIGIN + ci * G.hCell;
    const int maxY = min(iniY + G.hCell + 6, G.maxBY);
    const bool rowSkip = iniY >= G.maxBY - 3;            // ORBextractor.cpp:793
    const int sx0 = ORB_DET_ORIGIN + j0 * G.wCell;
    const int sx1 = min(ORB_DET_ORIGIN + min(j0 + FAST_WARPS, G.nCols) * G.wCell + 6, G.maxBX);
    const int srows = maxY - iniY;
    const int bufx0 = sx0 + ORB_EDGE, ax0 = bufx0 & ~3, shift = bufx0 - ax0;
    if (!rowSkip && sx1 > sx0) {
        const int nw = (shift + (sx1 - sx0) + 3) >> 2;
        const u8* src = pyr + (size_t)slot * P.pyr_bytes + G.pyr_ofs + (size_t)(iniY + ORB_EDGE) * G.pitch + ax0;
        for (int i = threadIdx.x; i < nw * srows; i += FAST_WARPS * 32) {
            const int r = i / nw, c = i - r * nw;
            *reinterpret_cast<u32*>(strip + r * SP + 4 * c) = *reinterpret_cast<const u32*>(src + (size_t)r * G.pitch + 4 * c);
        }
    }
    __syncthreads();

    const int j = j0 + warp;
    if (j >= G.nCols) return;
    const int cell = ci * G.nCols + j;
    int* cnt_out = cellcnt + (size_t)slot * P.ncells + G.cell_ofs + cell;
    const int iniX = ORB_DET_ORIGIN + j * G.wCell;
    const int maxX = min(iniX + G.wCell + 6, G.maxBX);
    const int cw = maxX - iniX - 6, ch = maxY - iniY - 6;   // detection window (FAST skips a 3-px rim)
    if (rowSkip || iniX >= G.maxBX - 6 || cw <= 0 || ch <= 0) {   // ORBextractor.cpp:793,801 / image < 7 px
        if (lane == 0) *cnt_out = 0;
        return;
    }
    u8* tile = smem + SP * SR + warp * (TP * TR);
    // zero the (ch+2) x (cw+2) score tile: the rim is what "outside the window scores 0" means
    for (int i = lane; i < (ch + 2) * TP; i += 32) tile[i] = 0;
    __syncwarp();
    const u8* s0 = strip + 3 * SP + (iniX - sx0) + shift + 3;
    for (int y = 0; y < ch; ++y)
        for (int x = lane; x < cw; x += 32) {
            const int s = fast_score(s0 + y * SP + x, SP, tLow);
            if (s) tile[(y + 1) * TP + x + 1] = (u8)s;
        }
    __syncwarp();
    const int iniTh = max(0, min(P.iniTh, 255)), minTh = max(0, min(P.minTh, 255));
    // pass 1: does anything survive NMS at iniThFAST?
    bool any_ini = false;
    for (int y = 0; y < ch; ++y)
        for (int x = lane; x < cw; x += 32) {
            const u8* t = tile + (y + 1) * TP + x + 1;
            const int s = t[0];
            if (s >= iniTh && s > 0 && s > t[-1] && s > t[1] && s > t[-TP - 1] && s > t[-TP] && s > t[-TP + 1] &&
                s > t[TP - 1] && s > t[TP] && s > t[TP + 1])
                any_ini = true;
        }
    const int T = __any_sync(0xffffffffu, any_ini) ? iniTh : minTh;
    // pass 2: ordered emission
    u32* out = cand + (size_t)slot * P.cand_entries + G.cand_ofs + (size_t)cell * G.cell_cap;
    int count = 0;
    const int xrel0 = iniX - ORB_DET_ORIGIN + 3, yrel0 = iniY - ORB_DET_ORIGIN + 3;
    for (int y = 0; y < ch; ++y)
        for (int xb = 0; xb < cw; xb += 32) {
            const int x = xb + lane;
            bool keep = false;
            int s = 0;
            if (x < cw) {
                const u8* t = tile + (y + 1) * TP + x + 1;
                s = t[0];
                keep = s >= T && s > 0 && s > t[-1] && s > t[1] && s > t[-TP - 1] && s > t[-TP] && s > t[-TP + 1] &&
                       s > t[TP - 1] && s > t[TP] && s > t[TP + 1];
            }
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[count + __popc(m & ((1u << lane) - 1))] = (u32)(xrel0 + x) | ((u32)(yrel0 + y) << 12) | ((u32)s << 24);
            count += __popc(m);
        }
    if (lane == 0) *cnt_out = count;
}
