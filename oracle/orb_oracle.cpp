// TEST INFRASTRUCTURE (oracle) -- never linked into or called from the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// CPU restatement of the reference stereo front-end:
//   * ORB extractor  : /root/reference/pyORBExtractor/ORBextractor.cpp:72-147, 410-852, 1033-1132
//   * stereo matching: /root/reference/Frame.py:161-279, 324-326 (NumPy >= 2 scalar semantics)
// Canonical choices where the reference is not reproducible against itself:
//   * DistributeOctTree ties between equal-sized nodes are broken by creation order (what the
//     reference does under a monotone allocator): the most recently created node is divided first.
//   * no FMA contraction (build with -ffp-contract=off).
// Parity pin: primitives vs cv2 4.13.0 (tests/test_oracle_prims.py); whole extractor vs the
// UNMODIFIED reference ORBextractor.cpp compiled against oracle/cvshim (oracle/_ref, built by
// oracle/Makefile); stereo vs fixtures produced by the reference's own Frame.compute_stereo_matches
// (tests/golden/make_golden.py).
#include "cvprims.hpp"
#include "../include/b200orb_pattern31.h"
#include <list>
#include <cstdio>
#include <thread>

using namespace orbo;

namespace {

const int kEdge = 19;       // EDGE_THRESHOLD  (ORBextractor.cpp:74)
const int kHalfPatch = 15;  // HALF_PATCH_SIZE (ORBextractor.cpp:73)
const int kPatch = 31;      // PATCH_SIZE      (ORBextractor.cpp:72)

static const signed char kPattern[1024] = {B200ORB_PATTERN_VALUES};

struct Kp { float x, y, size, angle, response; int octave; };

struct Level {
    int w = 0, h = 0;
    std::vector<u8> bordered;  // (h+38) x (w+38)
    size_t pitch() const { return (size_t)w + 2 * kEdge; }
    const u8* roi() const { return bordered.data() + (size_t)kEdge * pitch() + kEdge; }
    u8* roi() { return bordered.data() + (size_t)kEdge * pitch() + kEdge; }
};

struct Cand { int x, y, resp; };  // coordinates relative to the (16,16) detection origin

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;  // stored as double like ORBextractor.h:97
    std::vector<float> sf, isf, sig2, isig2;
    std::vector<int> quota, umax;
    std::vector<Level> pyr;
    std::vector<std::vector<Cand>> cands;   // last call, per level (FAST candidates fed to the tree)
    std::vector<std::vector<Kp>> levelKps;  // last call, per level, level coordinates (pre-scaling)
    std::vector<std::vector<u8>> blurred;   // last call, per level, w x h
};

// ORBextractor.cpp:410-470
Extractor* make_extractor(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    Extractor* e = new Extractor;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniTh; e->minTh = minTh;
    e->scaleFactor = scaleFactor;
    e->sf.assign(nlevels, 1.f); e->sig2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        e->sf[i] = (float)(e->sf[i - 1] * e->scaleFactor);  // float * double member -> double -> float
        e->sig2[i] = e->sf[i] * e->sf[i];
    }
    e->isf.resize(nlevels); e->isig2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) { e->isf[i] = 1.0f / e->sf[i]; e->isig2[i] = 1.0f / e->sig2[i]; }
    e->quota.resize(nlevels);
    float factor = (float)(1.0f / e->scaleFactor);
    float per = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e->quota[l] = cv_round(per);
        sum += e->quota[l];
        per *= factor;
    }
    e->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    e->umax.assign(kHalfPatch + 1, 0);
    int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) e->umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->pyr.resize(nlevels); e->cands.resize(nlevels); e->levelKps.resize(nlevels); e->blurred.resize(nlevels);
    return e;
}

// ORBextractor.cpp:1106-1132
void compute_pyramid(Extractor* e, const u8* img, int H, int W) {
    for (int l = 0; l < e->nlevels; ++l) {
        float s = e->isf[l];
        Level& L = e->pyr[l];
        L.w = cv_round((float)W * s);
        L.h = cv_round((float)H * s);
        L.bordered.assign((size_t)(L.w + 2 * kEdge) * (L.h + 2 * kEdge), 0);
        if (l == 0) {
            copy_make_border_101(img, W, H, W, L.bordered.data(), L.pitch(), kEdge);
        } else {
            const Level& P = e->pyr[l - 1];
            resize_linear_u8(P.roi(), P.w, P.h, P.pitch(), L.roi(), L.w, L.h, L.pitch());
            copy_make_border_101(L.roi(), L.w, L.h, L.pitch(), L.bordered.data(), L.pitch(), kEdge);
        }
    }
}

// ORBextractor.cpp:764-828 -- per-cell FAST with the iniThFAST -> minThFAST retry
void detect_cells(const Extractor* e, int l, std::vector<Cand>& out) {
    out.clear();
    const Level& L = e->pyr[l];
    const int minBX = kEdge - 3, minBY = minBX, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    if (nCols <= 0 || nRows <= 0) return;  // reference loops do not execute
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastKp> cell;
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const u8* sub = L.roi() + (ptrdiff_t)y0 * (ptrdiff_t)L.pitch() + x0;
            fast9_16_nms(sub, x1 - x0, y1 - y0, L.pitch(), e->iniTh, cell);
            if (cell.empty()) fast9_16_nms(sub, x1 - x0, y1 - y0, L.pitch(), e->minTh, cell);
            for (const FastKp& k : cell) out.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
        }
    }
}

// ---- DistributeOctTree restated (ORBextractor.cpp:481-762) ----
struct Node {
    int ulx, uly, brx, bry;   // UL and BR corners (UR.x == BR.x, BL.y == BR.y)
    std::vector<Cand> keys;
    bool leaf = false;        // bNoMore
    int seq = 0;              // creation order inside the current record (tie rule)
    std::list<Node>::iterator self;
};

void split4(const Node& n, Node c[4]) {  // DivideNode, ORBextractor.cpp:481-537
    const int hx = (int)std::ceil((float)(n.brx - n.ulx) / 2), hy = (int)std::ceil((float)(n.bry - n.uly) / 2);
    const int mx = n.ulx + hx, my = n.uly + hy;
    c[0] = Node{n.ulx, n.uly, mx, my, {}, false, 0, {}};
    c[1] = Node{mx, n.uly, n.brx, my, {}, false, 0, {}};
    c[2] = Node{n.ulx, my, mx, n.bry, {}, false, 0, {}};
    c[3] = Node{mx, my, n.brx, n.bry, {}, false, 0, {}};
    for (const Cand& k : n.keys) {
        int q = ((float)k.x < (float)mx) ? (((float)k.y < (float)my) ? 0 : 2) : (((float)k.y < (float)my) ? 1 : 3);
        c[q].keys.push_back(k);
    }
    for (int q = 0; q < 4; ++q) c[q].leaf = (c[q].keys.size() == 1);
}

void distribute(const std::vector<Cand>& in, int minX, int maxX, int minY, int maxY, int N,
                std::vector<Cand>& out) {
    out.clear();
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return;  // the reference divides by zero here (undefined); callers reject such shapes
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> L;
    std::vector<Node*> roots(std::max(nIni, 0));
    for (int i = 0; i < nIni; ++i) {
        Node n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY, {}, false, 0, {}};
        L.push_back(n);
        roots[i] = &L.back();
    }
    for (const Cand& k : in) roots[(int)((float)k.x / hX)]->keys.push_back(k);
    for (auto it = L.begin(); it != L.end();) {
        if (it->keys.empty()) it = L.erase(it);
        else { it->leaf = (it->keys.size() == 1); ++it; }
    }
    struct Rec { int size; int seq; Node* n; };
    std::vector<Rec> rec;
    auto push_children = [&](Node c[4], int& nToExpand) {
        for (int q = 0; q < 4; ++q) {
            if (c[q].keys.empty()) continue;
            L.push_front(c[q]);
            L.front().self = L.begin();
            if (c[q].keys.size() > 1) {
                ++nToExpand;
                L.front().seq = (int)rec.size();
                rec.push_back({(int)c[q].keys.size(), (int)rec.size(), &L.front()});
            }
        }
    };
    bool done = false;
    while (!done) {
        int prev = (int)L.size(), nToExpand = 0;
        rec.clear();
        for (auto it = L.begin(); it != L.end();) {
            if (it->leaf) { ++it; continue; }
            Node c[4];
            split4(*it, c);
            push_children(c, nToExpand);
            it = L.erase(it);
        }
        if ((int)L.size() >= N || (int)L.size() == prev) {
            done = true;
        } else if ((int)L.size() + nToExpand * 3 > N) {
            while (!done) {
                prev = (int)L.size();
                std::vector<Rec> order = rec;
                rec.clear();
                // ascending (size, creation order) -- creation order stands in for the node address
                std::sort(order.begin(), order.end(), [](const Rec& a, const Rec& b) {
                    return a.size != b.size ? a.size < b.size : a.seq < b.seq;
                });
                for (int j = (int)order.size() - 1; j >= 0; --j) {
                    Node c[4];
                    split4(*order[j].n, c);
                    int dummy = 0;
                    push_children(c, dummy);
                    L.erase(order[j].n->self);
                    if ((int)L.size() >= N) break;
                }
                if ((int)L.size() >= N || (int)L.size() == prev) done = true;
            }
        }
    }
    for (const Node& n : L) {
        const Cand* best = &n.keys[0];
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].resp > best->resp) best = &n.keys[k];
        out.push_back(*best);
    }
}

// IC_Angle, ORBextractor.cpp:77-104
float ic_angle(const Level& L, int cx, int cy, const std::vector<int>& umax) {
    const ptrdiff_t step = (ptrdiff_t)L.pitch();
    const u8* c = L.roi() + cy * step + cx;
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int p = c[u + v * step], m = c[u - v * step];
            vs += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vs;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// computeOrbDescriptor, ORBextractor.cpp:107-147 (no contraction; glibc sinf/cosf)
void describe(const u8* blur, int w, float px, float py, float angleDeg, u8* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float ang = angleDeg * factorPI, a, b;
    glibc_sincosf(ang, &b, &a);
    const u8* c = blur + (ptrdiff_t)cv_round(py) * w + cv_round(px);
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const signed char* p = kPattern + (i * 16 + k * 2) * 2;
            float x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            int t0 = c[(ptrdiff_t)cv_round(x0 * b + y0 * a) * w + cv_round(x0 * a - y0 * b)];
            int t1 = c[(ptrdiff_t)cv_round(x1 * b + y1 * a) * w + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (u8)val;
    }
}

// operator(), ORBextractor.cpp:1042-1104
int extract(Extractor* e, const u8* img, int H, int W, std::vector<Kp>& kps, std::vector<u8>& desc) {
    kps.clear(); desc.clear();
    if (!img || H <= 0 || W <= 0) return 0;
    compute_pyramid(e, img, H, W);
    for (int l = 0; l < e->nlevels; ++l) {
        const Level& L = e->pyr[l];
        detect_cells(e, l, e->cands[l]);
        std::vector<Cand> kept;
        distribute(e->cands[l], kEdge - 3, L.w - kEdge + 3, kEdge - 3, L.h - kEdge + 3, e->quota[l], kept);
        const int psize = (int)(kPatch * e->sf[l]);
        e->levelKps[l].clear();
        for (const Cand& k : kept)
            e->levelKps[l].push_back({(float)(k.x + kEdge - 3), (float)(k.y + kEdge - 3), (float)psize, -1.f,
                                      (float)k.resp, l});
    }
    for (int l = 0; l < e->nlevels; ++l)
        for (Kp& k : e->levelKps[l]) k.angle = ic_angle(e->pyr[l], cv_round(k.x), cv_round(k.y), e->umax);
    for (int l = 0; l < e->nlevels; ++l) {
        std::vector<Kp>& lk = e->levelKps[l];
        e->blurred[l].clear();
        if (lk.empty()) continue;
        const Level& L = e->pyr[l];
        // clone() of the level ROI, then blur with reflect-101 of the ROI itself
        std::vector<u8> roi((size_t)L.w * L.h);
        for (int y = 0; y < L.h; ++y) std::memcpy(&roi[(size_t)y * L.w], L.roi() + (size_t)y * L.pitch(), L.w);
        e->blurred[l].resize(roi.size());
        gaussian_blur7_u8(roi.data(), L.w, L.h, L.w, e->blurred[l].data(), L.w);
        for (const Kp& k : lk) {
            size_t off = desc.size();
            desc.resize(off + 32);
            describe(e->blurred[l].data(), L.w, k.x, k.y, k.angle, &desc[off]);
            Kp o = k;
            if (l != 0) { o.x = k.x * e->sf[l]; o.y = k.y * e->sf[l]; }
            kps.push_back(o);
        }
    }
    return (int)kps.size();
}

inline int popcount256(const u8* a, const u8* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

}  // namespace

extern "C" {

// ---------- primitives, exposed for the cv2 pin tests ----------
void orbo_resize(const u8* src, int sw, int sh, u8* dst, int dw, int dh) { resize_linear_u8(src, sw, sh, sw, dst, dw, dh, dw); }
void orbo_border101(const u8* src, int w, int h, u8* dst, int b) { copy_make_border_101(src, w, h, w, dst, w + 2 * b, b); }
void orbo_blur7(const u8* src, int w, int h, u8* dst) { gaussian_blur7_u8(src, w, h, w, dst, w); }
int orbo_fast(const u8* img, int w, int h, int threshold, int cap, int* xys /*[cap][3]*/) {
    std::vector<FastKp> out;
    fast9_16_nms(img, w, h, w, threshold, out);
    int n = (int)std::min<size_t>(out.size(), cap);
    for (int i = 0; i < n; ++i) { xys[3 * i] = out[i].x; xys[3 * i + 1] = out[i].y; xys[3 * i + 2] = out[i].score; }
    return (int)out.size();
}
void orbo_atan2(const float* y, const float* x, float* out, int n) { for (int i = 0; i < n; ++i) out[i] = fast_atan2_deg(y[i], x[i]); }
void orbo_sincos(const float* a, float* s, float* c, int n) { for (int i = 0; i < n; ++i) glibc_sincosf(a[i], s + i, c + i); }
// count of floats in [0, 2*pi] where the restated sincosf differs from the host libm (0 expected)
long orbo_sincos_exhaustive_mismatches(void) {
    const uint32_t hi = f2u(6.2831855f);
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<long> bad(nt, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            long mine = 0;
            for (uint64_t u = t; u <= hi; u += nt) {
                uint32_t b = (uint32_t)u;
                float y, s, c;
                std::memcpy(&y, &b, 4);
                glibc_sincosf(y, &s, &c);
                mine += (f2u(s) != f2u(sinf(y))) + (f2u(c) != f2u(cosf(y)));
            }
            bad[t] = mine;
        });
    for (auto& x : th) x.join();
    long total = 0;
    for (long b : bad) total += b;
    return total;
}
// standalone octree distribution: cand [n][3] = (x, y, resp) relative to the detection origin
int orbo_distribute(const int* cand, int n, int minX, int maxX, int minY, int maxY, int N, int cap, int* out) {
    std::vector<Cand> in(n), kept;
    for (int i = 0; i < n; ++i) in[i] = {cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]};
    distribute(in, minX, maxX, minY, maxY, N, kept);
    int m = (int)std::min<size_t>(kept.size(), cap);
    for (int i = 0; i < m; ++i) { out[3 * i] = kept[i].x; out[3 * i + 1] = kept[i].y; out[3 * i + 2] = kept[i].resp; }
    return (int)kept.size();
}

// ---------- extractor object ----------
void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    if (nfeatures < 0 || nlevels < 1 || !(scaleFactor > 1.f)) return nullptr;
    return make_extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbo_destroy(void* h) { delete (Extractor*)h; }
void orbo_tables(void* h, float* sf, float* isf, float* sig2, float* isig2, int* quota, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int l = 0; l < e->nlevels; ++l) { sf[l] = e->sf[l]; isf[l] = e->isf[l]; sig2[l] = e->sig2[l]; isig2[l] = e->isig2[l]; quota[l] = e->quota[l]; }
    for (int v = 0; v < 16; ++v) umax16[v] = e->umax[v];
}
// kps: [cap][6] floats (x, y, size, angle, response, octave); desc: [cap][32]; returns N (may exceed cap)
int orbo_extract(void* h, const u8* img, int H, int W, int cap, float* kps, u8* desc) {
    Extractor* e = (Extractor*)h;
    std::vector<Kp> k; std::vector<u8> d;
    int n = extract(e, img, H, W, k, d);
    int m = std::min(n, cap);
    for (int i = 0; i < m; ++i) {
        kps[6 * i] = k[i].x; kps[6 * i + 1] = k[i].y; kps[6 * i + 2] = k[i].size;
        kps[6 * i + 3] = k[i].angle; kps[6 * i + 4] = k[i].response; kps[6 * i + 5] = (float)k[i].octave;
    }
    if (m) std::memcpy(desc, d.data(), (size_t)m * 32);
    return n;
}
void orbo_level_size(void* h, int l, int* w, int* hh) { Extractor* e = (Extractor*)h; *w = e->pyr[l].w; *hh = e->pyr[l].h; }
void orbo_level_bordered(void* h, int l, u8* out) { Extractor* e = (Extractor*)h; std::memcpy(out, e->pyr[l].bordered.data(), e->pyr[l].bordered.size()); }
void orbo_level_blurred(void* h, int l, u8* out) { Extractor* e = (Extractor*)h; if (!e->blurred[l].empty()) std::memcpy(out, e->blurred[l].data(), e->blurred[l].size()); }
int orbo_level_candidates(void* h, int l, int cap, int* out) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->cands[l].size(), m = std::min(n, cap);
    for (int i = 0; i < m; ++i) { out[3 * i] = e->cands[l][i].x; out[3 * i + 1] = e->cands[l][i].y; out[3 * i + 2] = e->cands[l][i].resp; }
    return n;
}
int orbo_level_keypoints(void* h, int l, int cap, float* out /*[cap][4] x,y,resp,angle level coords*/) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->levelKps[l].size(), m = std::min(n, cap);
    for (int i = 0; i < m; ++i) { const Kp& k = e->levelKps[l][i]; out[4 * i] = k.x; out[4 * i + 1] = k.y; out[4 * i + 2] = k.response; out[4 * i + 3] = k.angle; }
    return n;
}
// The view Python gets from GetImagePyramid(): rows*cols CONTIGUOUS bytes from the ROI start of the
// bordered buffer (opencv_type_casters.h:205-240 ignores mat.step).
void orbo_level_caster_view(void* h, int l, u8* out) {
    Extractor* e = (Extractor*)h;
    const Level& L = e->pyr[l];
    std::memcpy(out, L.roi(), (size_t)L.w * L.h);
}

// ---------- stereo matching, Frame.py:161-279 under NumPy >= 2 scalar rules ----------
// kpsL/kpsR: [n][3] floats (x, y, octave).  pyrL/pyrR: per level pointer to the caster view (h_l x w_l).
// Outputs: uRight/depth (-1 = no match), bestIdx (Hamming winner or -1 if bestDist >= 75), bestDist.
// Returns 0, or -1 if a window would leave the view (the reference would raise there).
int orbo_stereo_ex(int NL, const float* kpsL, const u8* descL, int NR, const float* kpsR, const u8* descR,
                   int nlevels, const float* sf, const float* isf, const u8* const* pyrL, const u8* const* pyrR,
                   const int* lw, const int* lh, double mbf, float fx,
                   float* uRight, float* depth, int* bestIdx, int* bestDistOut, int* sadOut);
int orbo_stereo(int NL, const float* kpsL, const u8* descL, int NR, const float* kpsR, const u8* descR,
                int nlevels, const float* sf, const float* isf, const u8* const* pyrL, const u8* const* pyrR,
                const int* lw, const int* lh, double mbf, float fx,
                float* uRight, float* depth, int* bestIdx, int* bestDistOut) {
    return orbo_stereo_ex(NL, kpsL, descL, NR, kpsR, descR, nlevels, sf, isf, pyrL, pyrR, lw, lh, mbf, fx, uRight, depth, bestIdx,
                          bestDistOut, nullptr);
}
// sadOut (optional): the SAD minimum pushed to vDistIdx for accepted matches (Frame.py:279), -1 otherwise
int orbo_stereo_ex(int NL, const float* kpsL, const u8* descL, int NR, const float* kpsR, const u8* descR,
                   int nlevels, const float* sf, const float* isf, const u8* const* pyrL, const u8* const* pyrR,
                   const int* lw, const int* lh, double mbf, float fx,
                   float* uRight, float* depth, int* bestIdx, int* bestDistOut, int* sadOut) {
    const int nRows = lh[0];
    const float mbf32 = (float)mbf;
    const float mb = mbf32 / fx;       // Frame.py:43  (python float / np.float32 -> float32)
    const float maxD = mbf32 / mb;     // Frame.py:183
    std::vector<std::vector<int>> rows(nRows);
    for (int iR = 0; iR < NR; ++iR) {
        double y = kpsR[3 * iR + 1], r = 2.0 * (double)sf[(int)kpsR[3 * iR + 2]];
        int maxr = (int)std::ceil(y + r), minr = (int)std::floor(y - r);
        for (int yi = minr; yi <= maxr; ++yi) {
            if (yi < 0 || yi >= nRows) return -1;
            rows[yi].push_back(iR);
        }
    }
    int rc = 0;
    for (int iL = 0; iL < NL; ++iL) {
        uRight[iL] = -1.f; depth[iL] = -1.f;
        if (sadOut) sadOut[iL] = -1;
        if (bestIdx) bestIdx[iL] = -1;
        if (bestDistOut) bestDistOut[iL] = 100;
        const float uL = kpsL[3 * iL], vL = kpsL[3 * iL + 1];
        const int oL = (int)kpsL[3 * iL + 2];
        if ((int)vL < 0 || (int)vL >= nRows) { rc = -1; continue; }
        const std::vector<int>& cand = rows[(int)vL];
        if (cand.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL;
        if (maxU < 0) continue;
        int best = 100, bestR = 0;
        for (int iC : cand) {
            int oR = (int)kpsR[3 * iC + 2];
            if (oR < oL - 1 || oR > oL + 1) continue;
            float uR = kpsR[3 * iC];
            if (minU <= uR && uR <= maxU) {
                int d = popcount256(descL + 32 * (size_t)iL, descR + 32 * (size_t)iC);
                if (d < best) { best = d; bestR = iC; }
            }
        }
        if (bestDistOut) bestDistOut[iL] = best;
        if (!(best < 75.0)) continue;
        if (bestIdx) bestIdx[iL] = bestR;
        const double inv = (double)isf[oL];
        const long su = std::lrint((double)uL * inv), sv = std::lrint((double)vL * inv);
        const long sr = std::lrint((double)kpsR[3 * bestR] * inv);
        const int w = lw[oL], h = lh[oL];
        if (sr < 0 || sr + 11 >= w) continue;   // Frame.py:240-243
        if (sv - 5 < 0 || sv + 5 >= h || su - 5 < 0 || su + 5 >= w || sr - 10 < 0 || sr + 10 >= w) { rc = -1; continue; }
        const u8* A = pyrL[oL];
        const u8* B = pyrR[oL];
        int dist[11], bestSad = 0x7fffffff, bestInc = 0;
        const int lc = A[sv * w + su];
        for (int inc = -5; inc <= 5; ++inc) {
            const int rcv = B[sv * w + sr + inc];
            int s = 0;
            for (int dy = -5; dy <= 5; ++dy)
                for (int dx = -5; dx <= 5; ++dx)
                    s += std::abs((A[(sv + dy) * w + su + dx] - lc) - (B[(sv + dy) * w + sr + inc + dx] - rcv));
            dist[inc + 5] = s;
            if (s < bestSad) { bestSad = s; bestInc = inc; }
        }
        if (bestInc == -5 || bestInc == 5) continue;
        const float d1 = (float)dist[5 + bestInc - 1], d2 = (float)dist[5 + bestInc], d3 = (float)dist[5 + bestInc + 1];
        const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = sf[oL] * ((float)(sr + bestInc) + deltaR);
        float disparity = uL - bestuR;
        if (0 <= disparity && disparity < maxD) {
            if (sadOut) sadOut[iL] = bestSad;
            if (disparity <= 0) {   // Frame.py:273-275 (python-float branch)
                depth[iL] = (float)(mbf / 0.01);
                uRight[iL] = (float)((double)uL - 0.01);
            } else {
                depth[iL] = mbf32 / disparity;
                uRight[iL] = bestuR;
            }
        }
    }
    return rc;
}

}  // extern "C"
