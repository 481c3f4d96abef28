// TEST INFRASTRUCTURE (oracle) -- never linked into or called from the product path.
//
// CPU restatement of the OpenCV-owned arithmetic the reference extractor calls.  OpenCV 4.x is a
// third-party dependency of the reference (pyORBExtractor/CMakeLists.txt:15, unpinned; not vendored
// under /root/reference), so each primitive below restates OpenCV's published algorithm and is
// pinned bit-for-bit against cv2 4.13.0 by tests/test_oracle_prims.py (runs wherever cv2 imports).
//
// Call sites in the reference: resize ORBextractor.cpp:1120, copyMakeBorder :1122-1128,
// FAST :808-814, GaussianBlur :1085, fastAtan2 :103, cvRound/cvFloor/cvCeil :81,115,119-120,442,
// cosf/sinf :113 (glibc libm).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>

namespace orbo {

typedef unsigned char u8;

// ---- rounding helpers (OpenCV cvRound = round-half-to-even under the default FP environment) ----
static inline int cv_round(double v) { return (int)std::lrint(v); }
static inline int cv_round(float v) { return (int)std::lrintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }
static inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

// ---- BORDER_REFLECT_101 index map (OpenCV borderInterpolate) ----
static inline int reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

// dst (dh+2b) x (dw+2b), interior already holds / receives the src; fills every pixel from src with
// the reflect-101 map.  `src` may alias the interior of `dst` (the in-place BORDER_ISOLATED call).
static inline void copy_make_border_101(const u8* src, int w, int h, size_t sstep,
                                        u8* dst, size_t dstep, int b) {
    std::vector<int> xm(w + 2 * b);
    for (int x = 0; x < w + 2 * b; ++x) xm[x] = reflect101(x - b, w);
    // interior rows first (left/right borders), then top/bottom rows copied from finished rows
    for (int y = 0; y < h; ++y) {
        const u8* s = src + (size_t)y * sstep;
        u8* d = dst + (size_t)(y + b) * dstep;
        if (d + b != s) std::memmove(d + b, s, w);
        const u8* in = d + b;
        for (int x = 0; x < b; ++x) d[x] = in[xm[x]];
        for (int x = w + b; x < w + 2 * b; ++x) d[x] = in[xm[x]];
    }
    for (int y = 0; y < b; ++y)
        std::memcpy(dst + (size_t)y * dstep, dst + (size_t)(reflect101(y - b, h) + b) * dstep, w + 2 * b);
    for (int y = h + b; y < h + 2 * b; ++y)
        std::memcpy(dst + (size_t)y * dstep, dst + (size_t)(reflect101(y - b, h) + b) * dstep, w + 2 * b);
}

// ---- cv::resize INTER_LINEAR, CV_8UC1 (fixed-point path, INTER_RESIZE_COEF_BITS = 11) ----
struct ResizeTab {
    std::vector<int> xofs, y0, y1;       // source column; clamped source rows
    std::vector<short> xa0, xa1, yb0, yb1;
};
static inline ResizeTab make_resize_tab(int sw, int sh, int dw, int dh) {
    ResizeTab t;
    t.xofs.resize(dw); t.xa0.resize(dw); t.xa1.resize(dw);
    t.y0.resize(dh); t.y1.resize(dh); t.yb0.resize(dh); t.yb1.resize(dh);
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        t.xofs[dx] = sx;
        t.xa0[dx] = sat_short(cv_round((1.f - fx) * 2048.f));
        t.xa1[dx] = sat_short(cv_round(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        t.y0[dy] = std::min(std::max(sy, 0), sh - 1);
        t.y1[dy] = std::min(std::max(sy + 1, 0), sh - 1);
        t.yb0[dy] = sat_short(cv_round((1.f - fy) * 2048.f));
        t.yb1[dy] = sat_short(cv_round(fy * 2048.f));
    }
    return t;
}
static inline void resize_linear_u8(const u8* src, int sw, int sh, size_t sstep,
                                    u8* dst, int dw, int dh, size_t dstep) {
    ResizeTab t = make_resize_tab(sw, sh, dw, dh);
    for (int dy = 0; dy < dh; ++dy) {
        const u8* S0 = src + (size_t)t.y0[dy] * sstep;
        const u8* S1 = src + (size_t)t.y1[dy] * sstep;
        int b0 = t.yb0[dy], b1 = t.yb1[dy];
        u8* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = t.xofs[dx], sx1 = std::min(sx + 1, sw - 1);
            int a0 = t.xa0[dx], a1 = t.xa1[dx];
            int T0 = S0[sx] * a0 + S0[sx1] * a1;
            int T1 = S1[sx] * a0 + S1[sx1] * a1;
            D[dx] = (u8)((((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2);
        }
    }
}

// ---- cv::GaussianBlur 7x7 sigma=2 on CV_8U, BORDER_REFLECT_101 (fixed-point bit-exact path) ----
// Kernel in 8.8 fixed point = [18,34,48,56,48,34,18]; dst = (sum_y sum_x ky kx src + 2^15) >> 16.
static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};
static inline void gaussian_blur7_u8(const u8* src, int w, int h, size_t sstep, u8* dst, size_t dstep) {
    std::vector<unsigned> hbuf((size_t)w * h);
    std::vector<int> row(w + 6);
    for (int y = 0; y < h; ++y) {
        const u8* s = src + (size_t)y * sstep;
        for (int x = 0; x < w + 6; ++x) row[x] = s[reflect101(x - 3, w)];
        unsigned* hb = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; ++x)
            hb[x] = 18u * (row[x] + row[x + 6]) + 34u * (row[x + 1] + row[x + 5]) + 48u * (row[x + 2] + row[x + 4]) + 56u * row[x + 3];
    }
    for (int y = 0; y < h; ++y) {
        u8* d = dst + (size_t)y * dstep;
        const unsigned* r[7];
        for (int k = 0; k < 7; ++k) r[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        for (int x = 0; x < w; ++x) {
            unsigned acc = 18u * (r[0][x] + r[6][x]) + 34u * (r[1][x] + r[5][x]) + 48u * (r[2][x] + r[4][x]) + 56u * r[3][x];
            d[x] = (u8)((acc + 32768u) >> 16);
        }
    }
}

// ---- cv::FAST TYPE_9_16 with non-max suppression ----
static const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// threshold-free corner strength: largest t' such that the pixel is still a 9-of-16 corner at t'-1,
// i.e. max over the 16 arcs of min(v-p_k) and of min(p_k-v).  corner at t <=> best > t; score = best-1.
static inline int fast_best(const u8* p, size_t step) {
    int v = p[0], d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(ptrdiff_t)kFastDy[k] * (ptrdiff_t)step + kFastDx[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = 0;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

struct FastKp { int x, y, score; };
// 9-of-16 segment test at threshold t (strictly brighter than v+t or strictly darker than v-t).
// Antipodal quick rejects first (any 9-arc contains one pixel of every antipodal pair), then the run test.
static inline bool fast_is_corner(const u8* p, const ptrdiff_t* off, int t) {
    const int v = p[0], hi = v + t, lo = v - t;
    int a = p[off[0]], b = p[off[8]];
    int br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
    if (!(br | dk)) return false;
    a = p[off[4]]; b = p[off[12]];
    br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
    if (!(br | dk)) return false;
    unsigned mb = 0, md = 0;
    for (int k = 0; k < 16; ++k) { int q = p[off[k]]; mb |= (unsigned)(q > hi) << k; md |= (unsigned)(q < lo) << k; }
    for (unsigned m : {mb, md}) {
        m |= m << 16;                      // unroll the circle
        unsigned r = m & (m >> 1); r &= r >> 2; r &= r >> 4; r &= m >> 8;   // runs of 9
        if (r & 0xffffu) return true;
    }
    return false;
}
// FAST on a w x h sub-image; detections only in [3,w-3) x [3,h-3); NMS strict against the 8
// neighbours' scores (non-corners / outside the window count 0); row-major output.
static inline void fast9_16_nms(const u8* img, int w, int h, size_t step, int threshold,
                                std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    threshold = std::min(std::max(threshold, 0), 255);
    ptrdiff_t off[16];
    for (int k = 0; k < 16; ++k) off[k] = (ptrdiff_t)kFastDy[k] * (ptrdiff_t)step + kFastDx[k];
    std::vector<u8> sc((size_t)w * h, 0), is((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const u8* p = img + (size_t)y * step + x;
            if (!fast_is_corner(p, off, threshold)) continue;
            is[(size_t)y * w + x] = 1;
            sc[(size_t)y * w + x] = (u8)(fast_best(p, step) - 1);
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (!is[(size_t)y * w + x]) continue;
            const u8* s = &sc[(size_t)y * w + x];
            int v = s[0];
            if (v > s[-1] && v > s[1] && v > s[-w - 1] && v > s[-w] && v > s[-w + 1] &&
                v > s[w - 1] && v > s[w] && v > s[w + 1])
                out.push_back({x, y, v});
        }
}

// ---- cv::fastAtan2 (degrees), scalar path without FMA contraction ----
static inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- glibc 2.39 x86-64 sinf/cosf (ARM optimized-routines sincosf), restated; domain [0, 2*pi] ----
// All arithmetic in fp64, one rounding to fp32 at the end.  Pinned exhaustively against the host
// libm by tests (FMA contraction is immaterial).
struct SinCosTab { double sign[4]; double c0, c1, c2, c3, c4, s1, s2, s3; };
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline uint32_t abstop12(float f) { return (f2u(f) >> 20) & 0x7ff; }
static inline double sc_sin_poly(double x, double x2, int neg) {
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    (void)neg;
    double x3 = x * x2, s1p = s2 + x2 * s3, x7 = x3 * x2, s = x + x3 * s1;
    return s + x7 * s1p;
}
static inline double sc_cos_poly(double x2, int neg) {
    double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5,
           c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    if (neg) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; c4 = -c4; }
    double x4 = x2 * x2, c2p = c3 + x2 * c4, c1p = c1 + x2 * c2, x6 = x4 * x2, c = c0 + x2 * c1p;
    return c + x6 * c2p;
}
static inline void glibc_sincosf(float y, float* sp, float* cp) {
    double x = y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) { *sp = y; *cp = 1.0f; return; }
        *sp = (float)sc_sin_poly(x, x2, 0);
        *cp = (float)sc_cos_poly(x2, 0);
        return;
    }
    // reduction valid for |y| < 120 (the fast path of the routine); our domain is [0, 2*pi]
    double r = x * 0x1.45F306DC9C883p+23;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * 0x1.921FB54442D18p0;
    static const double sign[4] = {1.0, -1.0, -1.0, 1.0};
    double s = sign[n & 3];
    int tab = (n & 2) ? 1 : 0;
    double xs = x * s, x2 = x * x;
    // sin: polynomial chosen by n parity; cos: by (n^1) parity; table (negated cos constants) by n&2
    *sp = (float)((n & 1) ? sc_cos_poly(x2, tab) : sc_sin_poly(xs, x2, tab));
    *cp = (float)(((n ^ 1) & 1) ? sc_cos_poly(x2, tab) : sc_sin_poly(xs, x2, tab));
}

}  // namespace orbo
