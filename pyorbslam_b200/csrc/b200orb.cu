// b200orb: host side of the C ABI declared in include/b200orb.h.
// Builds the geometry plan, owns the device workspace, and enqueues the kernel sequence
//   K1 border + 7 x resize  ->  K5 blur  ->  K2 FAST cells  ->  K3 octree  ->  K4/K6 orient+describe  ->  K7/K8 stereo
// for S images at a time (S = 1 for the reference-compatible extractor object, 2 x pairs for the batch API).
// There is deliberately no CPU implementation behind these entry points.
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200orb.h"
#include "kernels_image.cuh"
#include "kernels_octree.cuh"
#include "kernels_features.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(B200ORB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
    } while (0)
#define TRY(expr) do { int _r = (expr); if (_r != 0) return _r; } while (0)

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Launch with programmatic stream serialization (see pdl_enter() in kernels_image.cuh): the kernel may start placing CTAs while the
// previous kernel of the stream drains; it blocks in griddepcontrol.wait until that kernel has completed.
// B200ORB_PDL: bit mask of the kernels launched that way (1 border0, 2 resize, 4 blur, 8 FAST, 16 octree, 32 describe, 64 row index, 128 stereo)
int g_pdl = [] { const char* v = getenv("B200ORB_PDL"); return v ? atoi(v) : 2; }();
template <typename... KArgs, typename... Args>
cudaError_t launch_seq(int pdl_bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (g_pdl & pdl_bit) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
inline int cv_round_f(float v) { return (int)lrintf(v); }

// ------------------------------------------------------------------------------------------------
// extractor constants: ORBextractor::ORBextractor, ORBextractor.cpp:410-470
// ------------------------------------------------------------------------------------------------
struct Params {
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    float scaleFactorF = 0;
    std::vector<float> sf, isf, sig2, isig2;
    std::vector<int> quota;
    int umax[16];
};

int make_params(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, Params& p) {
    if (nlevels < 1 || nlevels > ORB_MAX_LEVELS) return fail(B200ORB_E_ARG, "nlevels must be in [1,16]");
    if (nfeatures < 0) return fail(B200ORB_E_ARG, "nfeatures must be >= 0");
    if (!(scaleFactor > 1.0f)) return fail(B200ORB_E_ARG, "scaleFactor must be > 1");
    p.nfeatures = nfeatures; p.nlevels = nlevels; p.iniTh = iniTh; p.minTh = minTh; p.scaleFactorF = scaleFactor;
    const double scaleD = scaleFactor;                 // member `double scaleFactor`, ORBextractor.h:97
    p.sf.assign(nlevels, 1.f); p.sig2.assign(nlevels, 1.f); p.isf.resize(nlevels); p.isig2.resize(nlevels);
    for (int i = 1; i < nlevels; ++i) {
        p.sf[i] = (float)(p.sf[i - 1] * scaleD);
        p.sig2[i] = p.sf[i] * p.sf[i];
    }
    for (int i = 0; i < nlevels; ++i) { p.isf[i] = 1.0f / p.sf[i]; p.isig2[i] = 1.0f / p.sig2[i]; }
    p.quota.resize(nlevels);
    const float factor = (float)(1.0f / scaleD);
    float per = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        p.quota[l] = cv_round_f(per);
        sum += p.quota[l];
        per *= factor;
    }
    p.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax: end of each row of the radius-15 disc, made symmetric (ORBextractor.cpp:454-469)
    int um[17] = {0};
    const int vmax = (int)floor(15 * sqrtf(2.f) / 2 + 1), vmin = (int)ceil(15 * sqrtf(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) um[v] = (int)lrint(sqrt(225.0 - v * v));
    for (int v = 15, v0 = 0; v >= vmin; --v) {
        while (um[v0] == um[v0 + 1]) ++v0;
        um[v] = v0;
        ++v0;
    }
    for (int v = 0; v < 16; ++v) p.umax[v] = um[v];
    return 0;
}

// ------------------------------------------------------------------------------------------------
// geometry plan
// ------------------------------------------------------------------------------------------------
struct HostPlan {
    Plan P;
    std::vector<XTab> xtab;
    std::vector<XGroup> xgrp;
    std::vector<uint2> mtab;      // IC_Angle coefficient table, see k_describe
    std::vector<YTab> ytab;
    std::vector<unsigned char> roottab;   // k_octree: root index of every candidate column, per level
    std::vector<uint4> celltab;           // k_fast_cells: geometry of every cell of one image (see fast_cell_geom)
    int rs_gA_lo[ORB_MAX_LEVELS] = {0}, rs_gA_n[ORB_MAX_LEVELS] = {0}, rs_gB_n[ORB_MAX_LEVELS] = {0};   // k_resize: interior / border column groups
    int fast_SP = 0, fast_SR = 0, fast_TP = 0, fast_TR = 0, fast_LC = 0, fast_RQ = 0, fast_cells = 0, fast_WS = 0;
    LevelMaps maps;               // TMA tensor maps of the pyramid levels (k_fast_cells), rebuilt by Engine::plan
    LevelMaps bmaps;              // ... and of the blurred levels (k_describe: 64 x 37 boxes)
    size_t fast_smem = 0;
    int oct_capN = 0, oct_capK = 0, oct_capC = 0;
    size_t oct_smem = 0, oct_node_stride = 0;
    bool oct_global_nodes = false;
};

inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

int build_plan(const Params& prm, int H, int W, HostPlan& hp) {
    if (H < 1 || W < 1) return fail(B200ORB_E_ARG, "empty image");
    if (W > 4096 || H > 4096) return fail(B200ORB_E_ARG, "image larger than 4096 px per side is not supported (12-bit coordinates)");
    Plan& P = hp.P;
    memset(&P, 0, sizeof(P));
    P.nlevels = prm.nlevels; P.H = H; P.W = W; P.iniTh = prm.iniTh; P.minTh = prm.minTh;
    for (int v = 0; v < 16; ++v) P.umax[v] = prm.umax[v];
    hp.xtab.clear(); hp.ytab.clear(); hp.xgrp.clear(); hp.roottab.clear();
    int pyr = 0, blr = 0, cells = 0, cand = 0, kpt = 0, bctas = 0, maxw = 0, maxh = 0, maxcap = 0, maxcells = 0;
    for (int l = 0; l < prm.nlevels; ++l) {
        LevelGeom& G = P.lv[l];
        G.w = cv_round_f((float)W * prm.isf[l]);       // ORBextractor.cpp:1110-1111
        G.h = cv_round_f((float)H * prm.isf[l]);
        if (G.w < 1 || G.h < 1) return fail(B200ORB_E_ARG, "pyramid level collapses to zero size");
        G.pitch = round_up(G.w + 2 * ORB_EDGE, 64);
        G.rows = G.h + 2 * ORB_EDGE;
        G.pyr_ofs = pyr; pyr += round_up(G.pitch * G.rows, 256);
        G.blur_pitch = round_up(G.w, 64);
        // guard band in front of the first and behind the last level: the descriptor kernel stages the full 37 x 37 sample window
        // of every keypoint, which overhangs the level image by up to 2 rows for keypoints 16 px from its edge
        if (l == 0) blr = round_up(19 * G.blur_pitch + 64, 256);
        G.blur_ofs = blr; blr += round_up(G.blur_pitch * G.h, 256);
        if (l == P.nlevels - 1) blr += round_up(19 * G.blur_pitch + 64, 256);
        G.maxBX = G.w - ORB_EDGE + 3; G.maxBY = G.h - ORB_EDGE + 3;
        const float width = (float)(G.maxBX - ORB_DET_ORIGIN), height = (float)(G.maxBY - ORB_DET_ORIGIN);
        int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);    // ORBextractor.cpp:783-784
        if (nCols <= 0 || nRows <= 0) { nCols = 0; nRows = 0; }          // the reference's cell loops do not run
        G.nCols = nCols; G.nRows = nRows;
        if (nRows) { G.wCell = (int)ceilf(width / nCols); G.hCell = (int)ceilf(height / nRows); }
        G.cell_ofs = cells; cells += nRows * nCols;
        G.cell_cap = nRows ? ((G.wCell + 1) / 2) * ((G.hCell + 1) / 2) : 0;   // strict 8-neighbour maxima: <= 1 per 2x2 block
        G.cand_ofs = cand; cand += round_up(nRows * nCols * G.cell_cap, 4);
        G.blur_tiles_x = (G.w + BLUR_TW - 1) / BLUR_TW;
        G.blur_cta_ofs = bctas; bctas += G.blur_tiles_x * ((G.h + BLUR_TH - 1) / BLUR_TH);
        G.quota = prm.quota[l];
        G.nIni = 0; G.hX = 1.f;
        if (nRows) {
            G.nIni = (int)roundf((float)(G.maxBX - ORB_DET_ORIGIN) / (G.maxBY - ORB_DET_ORIGIN));   // ORBextractor.cpp:543
            if (G.nIni < 1) return fail(B200ORB_E_ARG, "image taller than 2x its width: the reference divides by zero here");
            G.hX = (float)(G.maxBX - ORB_DET_ORIGIN) / G.nIni;
            if (G.nIni > 250) return fail(B200ORB_E_ARG, "more than 250 octree roots (extreme aspect ratio)");
            G.root_ofs = (int)hp.roottab.size();
            // vpIniNodes[pt.x / hX] (ORBextractor.cpp:567): the float division evaluated here for every candidate column; the quotient
            // is clamped like nothing in the reference is -- x <= maxBX - 16 - 1 keeps it below nIni except by float rounding, where the
            // reference would index past its vector
            for (int x = 0; x <= 4095; ++x) {
                if (x > G.maxBX - ORB_DET_ORIGIN + 8) break;
                hp.roottab.push_back((unsigned char)std::min((int)((float)x / G.hX), G.nIni - 1));
            }
        }
        G.kp_cap = nRows ? std::max(4 * G.nIni, G.quota + 3) : 0;
        G.kp_ofs = kpt; kpt += G.kp_cap;
        G.sf = prm.sf[l]; G.isf = prm.isf[l];
        G.psize = (int)(31 * prm.sf[l]);
        maxw = std::max(maxw, G.wCell); maxh = std::max(maxh, G.hCell);
        maxcap = std::max(maxcap, G.kp_cap); maxcells = std::max(maxcells, nRows * nCols);
        if (l > 0) {   // cv::resize INTER_LINEAR tables, SURVEY.md App. A2
            const LevelGeom& S = P.lv[l - 1];
            const double sx_ = 1. / ((double)G.w / S.w), sy_ = 1. / ((double)G.h / S.h);
            std::vector<XTab> xr(G.w);
            std::vector<YTab> yr(G.h);
            for (int dx = 0; dx < G.w; ++dx) {
                float fx = (float)((dx + 0.5) * sx_ - 0.5);
                int sx = (int)floorf(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
                xr[dx] = XTab{sx, sat_short(cv_round_f((1.f - fx) * 2048.f)), sat_short(cv_round_f(fx * 2048.f))};
            }
            for (int dy = 0; dy < G.h; ++dy) {
                float fy = (float)((dy + 0.5) * sy_ - 0.5);
                int sy = (int)floorf(fy);
                fy -= sy;
                yr[dy] = YTab{(short)std::min(std::max(sy, 0), S.h - 1), (short)std::min(std::max(sy + 1, 0), S.h - 1),
                              sat_short(cv_round_f((1.f - fy) * 2048.f)), sat_short(cv_round_f(fy * 2048.f))};
            }
            // device tables are indexed by BORDERED coordinates: reflect-101 applied here, x padded to the pitch
            auto refl = [](int p, int len) { if (len == 1) return 0; while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p; return p; };
            while (hp.xtab.size() % 4) hp.xtab.push_back(XTab{0, 0, 0});      // 32-byte alignment of each level's row
            G.xtab_ofs = (int)hp.xtab.size(); G.ytab_ofs = (int)hp.ytab.size();
            const int bw = G.w + 2 * ORB_EDGE;
            for (int bx = 0; bx < G.pitch; ++bx) hp.xtab.push_back(xr[refl(std::min(bx, bw - 1) - ORB_EDGE, G.w)]);
            for (int by = 0; by < G.rows; ++by) hp.ytab.push_back(yr[refl(by - ORB_EDGE, G.h)]);
            // per-thread (4-column) entries of the word-window path; index = xtab_ofs / 4 + group
            hp.xgrp.resize(hp.xtab.size() / 4);
            for (int g = 0; g < G.pitch / 4; ++g) {
                XGroup e{0, 0xffffffffu, 0u, 0u, {0u, 0u, 0u, 0u}};
                const int xi = 4 * g - ORB_EDGE;
                bool ok = xi >= 0 && xi + 3 < G.w;
                for (int j = 0; ok && j < 4; ++j) { const int rel = xr[xi + j].sx - xr[xi].sx; ok = rel >= 0 && rel <= 6; }
                if (ok) {
                    const int c0 = xr[xi].sx + ORB_EDGE;
                    unsigned sel[4];
                    for (int j = 0; j < 4; ++j) {
                        const unsigned rel = (unsigned)(xr[xi + j].sx - xr[xi].sx);
                        sel[j] = rel | ((rel + 1) << 4) | (rel << 8) | (rel << 12);      // bytes 0 / 1 of the PRMT result = p[sx], p[sx + 1]
                        e.cf[j] = (unsigned)(unsigned short)xr[xi + j].a0 | ((unsigned)(unsigned short)xr[xi + j].a1 << 16);
                    }
                    e.wofs = c0 >> 2; e.shift8 = 8u * (unsigned)(c0 & 3);
                    e.sel01 = sel[0] | (sel[1] << 16); e.sel23 = sel[2] | (sel[3] << 16);
                }
                hp.xgrp[G.xtab_ofs / 4 + g] = e;
            }
            // interior groups (word-window path) must form one contiguous run; everything else with a valid column is a border group.
            // The word-window path needs the 4 columns of a thread to span <= 10 source bytes: scale < 2.
            const int gv = (bw + 3) / 4;
            int lo = -1, hi = -1;
            bool contiguous = (double)S.w / G.w < 1.95;
            for (int g = 0; g < gv && contiguous; ++g)
                if (hp.xgrp[G.xtab_ofs / 4 + g].shift8 != 0xffffffffu) { if (lo < 0) lo = g; else if (hi != g) contiguous = false; hi = g + 1; }
            if (!contiguous || lo < 0) { lo = 0; hi = 0; }
            hp.rs_gA_lo[l] = lo; hp.rs_gA_n[l] = hi - lo; hp.rs_gB_n[l] = gv - (hi - lo);
        }
    }
    // IC_Angle (ORBextractor.cpp:77-106) as dot products: the 31 x 31 disc is read as aligned 32-bit words, 3 rows x 9 words per
    // step (lane = (row % 3) * 9 + word), 11 steps; entry [al][step][lane] holds the 4 signed u coefficients and the 4 signed v
    // coefficients of that word's bytes (0 outside the disc), al = alignment of the patch's left edge.
    hp.mtab.assign(4 * MOM_STEPS * 32, make_uint2(0u, 0u));
    for (int al = 0; al < 4; ++al)
        for (int i = 0; i < MOM_STEPS; ++i)
            for (int lane = 0; lane < 32; ++lane) {
                const int k = lane % 9, row = 3 * i + lane / 9, v = row - 15;
                if (lane >= 27 || row > 30) continue;
                unsigned cu = 0, cv = 0;
                for (int j = 0; j < 4; ++j) {
                    const int u = 4 * k + j - al - 15;
                    if (std::abs(u) <= prm.umax[std::abs(v)]) { cu |= (unsigned)(u & 0xff) << (8 * j); cv |= (unsigned)(v & 0xff) << (8 * j); }
                }
                hp.mtab[(al * MOM_STEPS + i) * 32 + lane] = make_uint2(cu, cv);
            }
    P.pyr_bytes = pyr; P.blur_bytes = blr; P.ncells = std::max(cells, 1); P.cand_entries = std::max(cand, 4);
    P.kp_total = std::max(kpt, 1); P.blur_ctas = bctas; P.max_cells_level = maxcells;
    // the FAST kernel counts columns from the 4-byte boundary at or below the window's first pixel: up to 3 more than the cell is wide
    const int maxwa = maxw + 3;
    hp.fast_SP = round_up(15 + maxw + 6, 16);  // TMA box: starts at the 16-byte boundary below the window, width a multiple of 16
    hp.fast_SR = maxh + 6;
    hp.fast_TP = round_up(maxwa + 2, 4);
    hp.fast_TR = round_up(maxh + 2, 4);       // TP * TR is a multiple of 16: the tiles are cleared with 128-bit stores
    // wCell = ceil(width / floor(width / 30)) <= 59 whenever the level has cells at all
    if (maxwa > 64 || maxh > 63) return fail(B200ORB_E_ARG, "FAST cell larger than 61 x 63 px");
    // work list: half the window's pixels (a cell with more passing pixels is walked in row bands), at least one 64-pixel row
    const char* lc_str = getenv("B200ORB_FAST_LC");           // tests: a small capacity forces the banded path
    const int lc_env = lc_str ? atoi(lc_str) : 0;
    hp.fast_LC = round_up(std::max(lc_env > 0 ? lc_env : maxw * maxh / 2, 64), 8);            // 16-byte multiple
    hp.fast_RQ = round_up(std::max(maxh, 1), 4) * (maxwa > 32 ? 16 : 8);   // per warp: rows x quads x (min | ini nibble pair in a byte)
    // cell table (ORBextractor.cpp:783-806): window origin, detection size, level and candidate offset of every cell of an image
    hp.celltab.assign(std::max(cells, 1), make_uint4(0u, 0u, 0u, 0u));
    for (int l = 0; l < P.nlevels; ++l) {
        const LevelGeom& G = P.lv[l];
        for (int ci = 0; ci < G.nRows; ++ci)
            for (int j = 0; j < G.nCols; ++j) {
                const int local = ci * G.nCols + j;
                const int iniY = ORB_DET_ORIGIN + ci * G.hCell, iniX = ORB_DET_ORIGIN + j * G.wCell;
                const int maxY = std::min(iniY + G.hCell + 6, G.maxBY), maxX = std::min(iniX + G.wCell + 6, G.maxBX);
                int cw = maxX - iniX - 6, ch = maxY - iniY - 6;              // detection window (FAST skips a 3-px rim)
                if (iniY >= G.maxBY - 3 || iniX >= G.maxBX - 6 || ch <= 0 || cw <= 0) { cw = 0; ch = 0; }   // :793,801 / image < 7 px
                if (iniX > 0xffff || iniY > 0xffff) return fail(B200ORB_E_ARG, "image too large for the FAST cell table");
                // phase 1 counts columns from the word boundary at or below the first detection pixel (box column ((iniX + ORB_EDGE) & 15) + 3)
                const int xa = (((iniX + ORB_EDGE) & 15) + 3) & 3, Q = std::max((cw + xa + 3) >> 2, 1);
                const unsigned rcp = (65536u + Q - 1) / Q, dr = 32u / Q;
                hp.celltab[G.cell_ofs + local] = make_uint4((unsigned)iniX | ((unsigned)iniY << 16), (unsigned)cw | ((unsigned)ch << 8) | ((unsigned)l << 16),
                                                            (unsigned)(G.cand_ofs + local * G.cell_cap), rcp | (dr << 17) | ((unsigned)Q << 23));
            }
    }
    hp.fast_WS = round_up(hp.fast_SP * hp.fast_SR + hp.fast_TP * hp.fast_TR + hp.fast_LC * 2 + hp.fast_RQ + 16, 128);   // +16: phase 1 reads whole words past the last row
    hp.fast_smem = (size_t)FAST_WARPS * hp.fast_WS;
    if (hp.fast_smem > 200 * 1024) return fail(B200ORB_E_ARG, "cell size too large for the FAST kernel's shared memory");
    hp.fast_cells = cells;
    hp.oct_capN = round_up(maxcap + 8, 4);
    hp.oct_capC = round_up(std::max(maxcells, 1), 4);
    // node arrays go to shared memory when they fit next to >= 2048 keys in ~200 KB, else to a global scratch block
    const size_t nodeB = oct_node_bytes(hp.oct_capN);
    hp.oct_global_nodes = nodeB + oct_base_bytes(2048, hp.oct_capC) > 200 * 1024;
    const size_t fixed = oct_base_bytes(0, hp.oct_capC) + (hp.oct_global_nodes ? 0 : nodeB);
    if (fixed + 1024 * 8 > 200 * 1024) return fail(B200ORB_E_ARG, "image has too many FAST cells per level for the octree kernel's shared memory");
    const char* room_kb = getenv("B200ORB_OCT_ROOM_KB");
    // shared memory per CTA: 74 KB (three CTAs per SM) if at least 4096 keys fit next to the fixed part, else 110 KB (two), else 200 KB (one)
    long long room = 0;
    for (int kb : {room_kb ? atoi(room_kb) : 74, 110, 200}) {
        room = (long long)kb * 1024 - (long long)fixed;
        if (room >= 4096 * 8) break;
    }
    hp.oct_capK = (int)std::min<long long>(room / 8, 16384) & ~3;
    hp.oct_node_stride = (nodeB + 255) & ~(size_t)255;
    hp.oct_smem = oct_base_bytes(hp.oct_capK, hp.oct_capC) + (hp.oct_global_nodes ? 0 : nodeB);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Engine: workspace for S image slots + the kernel sequence
// ------------------------------------------------------------------------------------------------
constexpr int kStatusBanks = 4;      // = b200orb_batch::NBUF
constexpr int kMaxDynSmem = 200 * 1024;   // upper bound of the dynamic shared memory any launch of k_fast_cells / k_octree asks for
struct Engine {
    Params prm;
    HostPlan hp;
    int device = 0, S = 0;
    bool planned = false;
    u8 *d_pyr = nullptr, *d_blur = nullptr;
    u32 *d_cand = nullptr, *d_scratch = nullptr, *d_lvlkp = nullptr;
    int *d_cellcnt = nullptr, *d_lvlcnt = nullptr, *d_status = nullptr, *d_rowstart = nullptr;
    uint2* d_rmeta = nullptr;     // row-band index entries of the right keypoints: band_rows() x kp_total per pair
    unsigned char* d_octnodes = nullptr;     // node arrays of k_octree when they do not fit in shared memory
    unsigned char* d_roottab = nullptr;
    uint4* d_celltab = nullptr;
    int* d_fastctr = nullptr;                // k_fast_cells: next unclaimed cell
    XTab* d_xtab = nullptr;
    XGroup* d_xgrp = nullptr;
    uint2* d_mtab = nullptr;
    float4* d_fpat = nullptr;
    YTab* d_ytab = nullptr;
    long long bytes = 0;
    // K5 (blur) depends only on the pyramid, K2 -> K3 (FAST -> octree) likewise: per launch sequence the blur is forked onto a side
    // stream so that it fills the SMs the latency-bound octree rounds leave idle; both join in front of K4/K6 (describe)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int fork_blur = -1;      // -1 = read B200ORB_FORK_BLUR (default off: measured, no gain -- the block scheduler drains the earlier launch first)
    int sm_count = 148;

    ~Engine() {
        if (side) { cudaSetDevice(device); cudaStreamDestroy(side); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); }
    }
    void release() {
        cudaFree(d_pyr); cudaFree(d_blur); cudaFree(d_cand); cudaFree(d_scratch); cudaFree(d_lvlkp);
        cudaFree(d_cellcnt); cudaFree(d_lvlcnt); cudaFree(d_status); cudaFree(d_xtab); cudaFree(d_xgrp); d_xgrp = nullptr; cudaFree(d_mtab); d_mtab = nullptr; cudaFree(d_fpat); d_fpat = nullptr; cudaFree(d_ytab); cudaFree(d_rowstart); cudaFree(d_rmeta); d_rmeta = nullptr; cudaFree(d_octnodes); d_octnodes = nullptr; cudaFree(d_roottab); d_roottab = nullptr; cudaFree(d_celltab); d_celltab = nullptr; cudaFree(d_fastctr); d_fastctr = nullptr;
        d_pyr = d_blur = nullptr; d_cand = d_scratch = d_lvlkp = nullptr; d_cellcnt = d_lvlcnt = d_status = d_rowstart = nullptr;
        d_xtab = nullptr; d_ytab = nullptr; planned = false; bytes = 0;
    }
    template <typename T> int alloc(T** p, size_t n) {
        CU_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
        bytes += (long long)(n * sizeof(T));
        return 0;
    }
    // one 3-D tensor map per level over the slots' bordered buffers: {x: pitch bytes, y: rows, z: slot}; box = the FAST window
    int make_level_maps() {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
            if (!fn || q != cudaDriverEntryPointSuccess) return fail(B200ORB_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
            encode = (EncodeFn)fn;
        }
        const Plan& P = hp.P;
        memset(&hp.maps, 0, sizeof(hp.maps));
        memset(&hp.bmaps, 0, sizeof(hp.bmaps));
        for (int l = 0; l < P.nlevels; ++l) {
            const LevelGeom& G = P.lv[l];
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            {   // bordered raw level, box = FAST window
                const cuuint64_t dims[3] = {(cuuint64_t)G.pitch, (cuuint64_t)G.rows, (cuuint64_t)S};
                const cuuint64_t strides[2] = {(cuuint64_t)G.pitch, (cuuint64_t)P.pyr_bytes};      // bytes, dimensions 1 and 2
                const cuuint32_t box[3] = {(cuuint32_t)hp.fast_SP, (cuuint32_t)hp.fast_SR, 1u};
                const CUresult r = encode(&hp.maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_pyr + G.pyr_ofs, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(B200ORB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
            }
            {   // blurred level (dense rows), box = the descriptor's 37-row sample window
                const cuuint64_t dims[3] = {(cuuint64_t)G.blur_pitch, (cuuint64_t)G.h, (cuuint64_t)S};
                const cuuint64_t strides[2] = {(cuuint64_t)G.blur_pitch, (cuuint64_t)P.blur_bytes};
                const cuuint32_t box[3] = {(cuuint32_t)DESC_PP, 37u, 1u};
                const CUresult r = encode(&hp.bmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_blur + G.blur_ofs, dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(B200ORB_E_CUDA, "cuTensorMapEncodeTiled (blur) failed with CUresult " + std::to_string((int)r));
            }
        }
        return 0;
    }
    int plan(int H, int W, int slots) {
        if (planned && hp.P.H == H && hp.P.W == W && S == slots) return 0;
        CU_TRY(cudaSetDevice(device));
        release();
        TRY(build_plan(prm, H, W, hp));
        S = slots;
        const Plan& P = hp.P;
        TRY(alloc(&d_pyr, (size_t)S * P.pyr_bytes));
        TRY(alloc(&d_blur, (size_t)S * P.blur_bytes));
        TRY(alloc(&d_cand, (size_t)S * P.cand_entries));
        TRY(alloc(&d_scratch, (size_t)S * P.cand_entries * 2));
        TRY(alloc(&d_lvlkp, (size_t)S * P.kp_total));
        TRY(alloc(&d_cellcnt, (size_t)S * P.ncells));
        TRY(alloc(&d_lvlcnt, (size_t)S * P.nlevels));
        TRY(alloc(&d_status, (size_t)kStatusBanks * std::max(S, 1)));     // one range-error flag word per pair, one bank per run_host chunk in flight
        TRY(alloc(&d_rowstart, (size_t)S * (P.lv[0].h + 1)));
        TRY(alloc(&d_rmeta, (size_t)std::max(S / 2, 1) * P.kp_total * band_rows()));
        if (hp.oct_global_nodes) TRY(alloc(&d_octnodes, (size_t)S * P.nlevels * hp.oct_node_stride));
        TRY(alloc(&d_fastctr, 1));
        TRY(alloc(&d_roottab, hp.roottab.size() + 16));
        if (!hp.roottab.empty()) CU_TRY(cudaMemcpy(d_roottab, hp.roottab.data(), hp.roottab.size(), cudaMemcpyHostToDevice));
        TRY(alloc(&d_celltab, hp.celltab.size()));
        CU_TRY(cudaMemcpy(d_celltab, hp.celltab.data(), hp.celltab.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        TRY(alloc(&d_xtab, hp.xtab.size()));
        TRY(alloc(&d_xgrp, hp.xgrp.size()));
        TRY(alloc(&d_mtab, hp.mtab.size()));
        TRY(alloc(&d_fpat, 256));
        TRY(alloc(&d_ytab, hp.ytab.size()));
        CU_TRY(cudaMemset(d_pyr, 0, (size_t)S * P.pyr_bytes));
        CU_TRY(cudaMemset(d_blur, 0, (size_t)S * P.blur_bytes));
        CU_TRY(cudaMemset(d_status, 0, sizeof(int) * (size_t)kStatusBanks * std::max(S, 1)));
        if (!hp.xtab.empty()) CU_TRY(cudaMemcpy(d_xtab, hp.xtab.data(), hp.xtab.size() * sizeof(XTab), cudaMemcpyHostToDevice));
        if (!hp.xgrp.empty()) CU_TRY(cudaMemcpy(d_xgrp, hp.xgrp.data(), hp.xgrp.size() * sizeof(XGroup), cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_mtab, hp.mtab.data(), hp.mtab.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        {
            static const signed char pat[1024] = {B200ORB_PATTERN_VALUES};
            std::vector<float4> fp(256);
            for (int lane = 0; lane < 32; ++lane)
                for (int k = 0; k < 8; ++k) {
                    const signed char* q = pat + 4 * (8 * lane + k);
                    fp[k * 32 + lane] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
                }
            CU_TRY(cudaMemcpy(d_fpat, fp.data(), fp.size() * sizeof(float4), cudaMemcpyHostToDevice));
        }
        if (!hp.ytab.empty()) CU_TRY(cudaMemcpy(d_ytab, hp.ytab.data(), hp.ytab.size() * sizeof(YTab), cudaMemcpyHostToDevice));
        TRY(make_level_maps());
        // the attribute is per kernel and device, not per engine: every engine sets the same upper bound (the most any geometry may ask
        // for), or an engine planned later with a smaller need would make the launches of an earlier one fail
        CU_TRY(cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        CU_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        if ((int)hp.oct_smem > kMaxDynSmem || (int)hp.fast_smem > kMaxDynSmem) return fail(B200ORB_E_ARG, "kernel shared memory beyond the per-CTA limit");
        CU_TRY(cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
        planned = true;
        return 0;
    }
    // images: slots [0, splitA) from imgA, [splitA, n) from imgB; outputs: kps [n][kp_total][6], desc [n][kp_total][32], nkp [n]
    // evs (optional): B200ORB_NSTAGE + 1 events; evs[0] must already be recorded by the caller, evs[k + 1] is
    // recorded after stage k (0 border, 1 resize chain, 2 blur, 3 FAST, 4 octree, 5 describe; 6 = stereo, by the caller)
    int extract(const u8* imgA, const u8* imgB, int splitA, int n, float* d_kps, u8* d_desc, int* d_nkp, cudaStream_t st,
                cudaEvent_t* evs = nullptr) {
        if (n < 1 || n > S) return fail(B200ORB_E_ARG, "slot count out of range");
        const Plan& P = hp.P;
        auto magic_of = [](int d) { return d > 1 ? (unsigned)((0x100000000ULL + (unsigned)d - 1) / (unsigned)d) : 0xffffffffu; };
        {
            // 16-byte spans per row: interior ones (source bytes x0 .. x0 + 19 inside the image row) and rim ones, in separate blocks
            const LevelGeom& G = P.lv[0];
            const int nv = (G.w + 2 * ORB_EDGE + 15) / 16;
            const int vA_lo = 2, vA_n = std::max(0, (P.W + 15) / 16 - vA_lo), vB_n = nv - vA_n;
            const int blkA = (vA_n * G.rows + 255) / 256, blkB = (vB_n * G.rows + 255) / 256;
            CU_TRY(launch_seq(1, k_border0, dim3(blkA + blkB, n), dim3(256), 0, st, P, imgA, imgB, splitA, d_pyr, blkA, vA_lo, vA_n, magic_of(vA_n), vB_n, magic_of(vB_n), d_fastctr));
            ++g_launches;
            if (evs) cudaEventRecord(evs[1], st);
        }
        // one launch per level over all images.  (Tried: the small upper levels in one launch, a CTA per image walking them with block
        // barriers in between -- 0.364 vs 0.367 ms per 128 pairs, not worth a second code path.)
        for (int l = 1; l < P.nlevels; ++l) {
            const LevelGeom& G = P.lv[l];
            const int rgroups = (G.rows + RS_ROWS - 1) / RS_ROWS;
            const int blkA = (hp.rs_gA_n[l] * rgroups + 255) / 256, blkB = (hp.rs_gB_n[l] * rgroups + 255) / 256;
            CU_TRY(launch_seq(2, k_resize, dim3(blkA + blkB, n), dim3(256), 0, st, P, l, blkA, hp.rs_gA_lo[l], hp.rs_gA_n[l], magic_of(hp.rs_gA_n[l]), hp.rs_gB_n[l],
                              magic_of(hp.rs_gB_n[l]), d_pyr, (const XTab*)d_xtab, (const XGroup*)d_xgrp, (const YTab*)d_ytab));
            ++g_launches;
        }
        if (evs) cudaEventRecord(evs[2], st);
        if (fork_blur < 0) { const char* v = getenv("B200ORB_FORK_BLUR"); fork_blur = v ? atoi(v) : 0; }
        const bool fork = fork_blur > 0 && !evs;      // the per-stage timing pass runs the stages back to back on one stream
        if (fork && !side) {
            int lo = 0, hi = 0;
            CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, fork_blur >= 3 ? lo : hi));
            CU_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            CU_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        auto launch_blur = [&]() {
            if (fork) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(side, ev_fork, 0); }
            launch_seq(4, k_blur, dim3(P.blur_ctas, n), dim3(BLUR_WARPS * 32), 0, fork ? side : st, P, (const u8*)d_pyr, d_blur);
            ++g_launches;
            if (fork) cudaEventRecord(ev_join, side);
        };
        if (!fork || fork_blur == 1) launch_blur();
        if (evs) cudaEventRecord(evs[3], st);
        if (hp.fast_cells > 0) {
            // persistent: every warp walks cells of the whole launch; 8 CTAs of FAST_WARPS warps per SM are resident
            const int total = n * hp.fast_cells;
            const int ctas = std::min((total + FAST_WARPS - 1) / FAST_WARPS, sm_count * 8);
            // the cell counter was zeroed by this sequence's k_border0
            CU_TRY(launch_seq(8, k_fast_cells, dim3(ctas), dim3(FAST_WARPS * 32), hp.fast_smem, st, P, hp.maps, (const uint4*)d_celltab, d_cand, d_cellcnt, total, hp.fast_cells, magic_of(hp.fast_cells), d_fastctr, 1u,
                              hp.fast_SP, hp.fast_SR, hp.fast_TP, hp.fast_TR, hp.fast_LC, hp.fast_RQ, hp.fast_WS));
            ++g_launches;
        }
        if (fork && fork_blur >= 2) launch_blur();
        if (evs) cudaEventRecord(evs[4], st);
        CU_TRY(launch_seq(16, k_octree, dim3(n, P.nlevels), dim3(OCT_THREADS), hp.oct_smem, st, P, (const u32*)d_cand, (const int*)d_cellcnt, d_scratch, d_lvlkp,
                          d_lvlcnt, hp.oct_capN, hp.oct_capK, hp.oct_capC, d_octnodes, hp.oct_node_stride, (const unsigned char*)d_roottab));
        ++g_launches;
        if (evs) cudaEventRecord(evs[5], st);
        if (fork) CU_TRY(cudaStreamWaitEvent(st, ev_join, 0));
        CU_TRY(launch_seq(32, k_describe, dim3((P.kp_total + DESC_WARPS * DESC_KPW - 1) / (DESC_WARPS * DESC_KPW), n), dim3(DESC_WARPS * 32), 0, st, P, hp.bmaps,
                          (const u8*)d_pyr, (const u32*)d_lvlkp, (const int*)d_lvlcnt, (const uint2*)d_mtab, (const float4*)d_fpat, d_kps, d_desc, d_nkp));
        ++g_launches;
        if (evs) cudaEventRecord(evs[6], st);
        CU_TRY(cudaGetLastError());
        return 0;
    }
    // upper bound of the rows one right keypoint is entered into: floor(y - 2s) .. ceil(y + 2s) at the coarsest level
    int band_rows() const {
        float smax = 1.f;
        for (float v : prm.sf) smax = std::max(smax, v);
        return 2 * (int)ceil(2.0 * smax) + 2;
    }
    // flags: B200ORB_STEREO_DENSE_PYRAMID -> SAD windows on the true level image instead of the reference's sheared view
    void stereo_geom(StereoGeom& SG, int flags = 0) const {
        const Plan& P = hp.P;
        memset(&SG, 0, sizeof(SG));
        SG.nlevels = P.nlevels;
        SG.nRows = P.lv[0].h;
        for (int l = 0; l < P.nlevels; ++l) {
            const LevelGeom& G = P.lv[l];
            SG.sf[l] = G.sf; SG.isf[l] = G.isf; SG.w[l] = G.w; SG.h[l] = G.h;
            SG.plog[l] = G.w + 2 * ORB_EDGE;                 // the reference's Mat step
            SG.pitch[l] = G.pitch;
            SG.magic[l] = (unsigned)((0x100000000ULL + SG.plog[l] - 1) / SG.plog[l]);
            SG.off0[l] = ORB_EDGE * SG.plog[l] + ORB_EDGE;
            SG.vstride[l] = (flags & B200ORB_STEREO_DENSE_PYRAMID) ? SG.plog[l] : G.w;
            SG.base[l] = G.pyr_ofs;
        }
    }
};

void fill_stereo_consts(StereoArgs& A, double mbf, float fx) {
    A.mbf = mbf;
    A.mbf32 = (float)mbf;            // python float / np.float32 -> float32 (NumPy >= 2), Frame.py:43
    A.mb = A.mbf32 / fx;
    A.maxD = A.mbf32 / A.mb;         // Frame.py:183
}

// row index of the right keypoints, then the matcher.  A.rowStart / A.rmeta / A.idx_stride must point at
// (nRows + 1) and idx_stride ints per pair of scratch.
int launch_stereo(const StereoGeom& SG, StereoArgs A, int max_left, int pairs, cudaStream_t st, int flags = 0) {
    if (pairs < 1 || max_left < 1) return 0;
    const size_t smem = (size_t)(2 * SG.nRows + 1) * sizeof(int);
    CU_TRY(launch_seq(64, k_rowindex, dim3(pairs), dim3(RI_THREADS), smem, st, A.kpsR, A.nR, A.kp_stride, A.n_stride, A.kp_row, A.oct_idx, SG, (int*)A.rowStart,
                      (uint2*)A.rmeta, A.idx_stride, A.status, A.status_stride));
    ++g_launches;
    dim3 grid((max_left + ST_WARPS - 1) / ST_WARPS, pairs);
    CU_TRY(launch_seq(128, k_stereo, grid, dim3(ST_WARPS * 32), 0, st, SG, A));
    ++g_launches;
    if (flags & B200ORB_STEREO_MEDIAN_CULL) {
        if (!A.sadDist) return fail(B200ORB_E_ARG, "median cull needs a sadDist buffer");
        CU_TRY(launch_seq(128, k_median_cull, dim3(pairs), dim3(MC_THREADS), 0, st, A.nL, A.n_stride, A.out_stride, (const int*)A.sadDist, A.uRight, A.depth));
        ++g_launches;
    }
    CU_TRY(cudaGetLastError());
    return 0;
}

// Grow-only device scratch per device for the stateless entry points (all-pairs Hamming, area queries): carving a cached block
// replaces a dozen cudaMalloc / cudaFree pairs per call.  Calls on one device are serialised by the arena's mutex.
struct ScratchArena {
    std::mutex mu;
    unsigned char* base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes) {
        used = 0;
        if (bytes <= cap) return 0;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        CU_TRY(cudaMalloc((void**)&base, bytes + bytes / 4 + 4096));
        cap = bytes + bytes / 4 + 4096;
        return 0;
    }
    template <typename T> T* take(size_t n) {
        unsigned char* p = base + used;
        used += (n * sizeof(T) + 255) & ~(size_t)255;
        return reinterpret_cast<T*>(p);
    }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};
ScratchArena& arena_of(int device) {
    static std::mutex m;
    static std::map<int, ScratchArena> arenas;
    std::lock_guard<std::mutex> g(m);
    return arenas[device];
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
struct b200orb_extractor {
    Engine eng;
    cudaStream_t st = nullptr, st_pyr = nullptr;    // st: the launch sequence; st_pyr: background download of the pyramid (GetImagePyramid)
    cudaEvent_t ev_done = nullptr, ev_pyr = nullptr;
    u8* d_img = nullptr; size_t img_cap = 0;
    // results of the last call in ONE device block [nkp, pad to 256 B | kps C x 24, pad to 256 B | desc C x 32] so that one copy brings them back
    u8* d_out = nullptr; u8* h_out = nullptr;        // device block / its pinned host mirror
    float* d_kps = nullptr; u8* d_desc = nullptr; int* d_nkp = nullptr;       // views into d_out
    // stereo outputs likewise: [uRight C | depth C | matchIdx C | sadDist C | status]
    u8* d_sout = nullptr; u8* h_sout = nullptr;
    float *d_uR = nullptr, *d_depth = nullptr; int *d_match = nullptr, *d_sad = nullptr, *d_sstatus = nullptr;
    int out_cap = 0;
    u8* h_img = nullptr; size_t h_img_cap = 0;       // pinned upload staging
    u8* h_pyr = nullptr; size_t h_pyr_cap = 0;       // pinned mirror of the pyramid blob (physical layout)
    bool pyr_inflight = false, pyr_valid = false;
    // the launch sequence of one image (upload, 8 pyramid kernels, blur, FAST, octree, describe, result download) as a CUDA graph,
    // captured once per image size: one cudaGraphLaunch instead of ~17 stream operations
    cudaGraphExec_t graph = nullptr;
    int graph_H = 0, graph_W = 0;
    long long graph_kernels = 0;     // kernels one graph launch runs (for b200orb_kernel_launches)
    int use_graph = -1;          // -1: read B200ORB_GRAPH (default 1)
    int prefetch_pyramid = 1;    // start the pyramid download behind the results (Frame.__init__ always asks for it, Frame.py:59-60)
    int n = -1;          // keypoints of the last call, -1 = none yet
    bool empty_last = false;
    void drop_graph() { if (graph) { cudaGraphExecDestroy(graph); graph = nullptr; } }
};

struct b200orb_vocab {
    int device = 0, n_nodes = 0;
    int *d_child_begin = nullptr, *d_child_ids = nullptr;
    u8* d_node_desc = nullptr;
    // scratch for transform calls
    u8* d_desc = nullptr; int *d_leaf = nullptr, *d_level = nullptr; int cap = 0;
};

struct b200orb_batch {
    Engine eng;
    int P = 0, H = 0, W = 0;
    // run_host pipeline state
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_comp2 = nullptr, s_out = nullptr;
    // run_host keeps NBUF chunks in flight (upload k+3 | compute k+2 and k+1 | download k).  Consecutive chunks alternate between two
    // compute lanes -- each its own stream and Engine workspace -- so that one chunk's latency-bound launches (octree rounds, the small
    // pyramid levels, kernel tails) run under the other chunk's kernels; B200ORB_HOST_LANES=1 keeps one lane
    static constexpr int NBUF = kStatusBanks;
    Engine* eng2 = nullptr;
    int lanes = 1;
    cudaEvent_t ev_in[NBUF] = {}, ev_comp[NBUF] = {}, ev_out[NBUF] = {};
    u8* d_in[NBUF] = {};
    float* d_kps[NBUF] = {}; u8* d_desc[NBUF] = {}; int* d_nkp[NBUF] = {};
    float* d_uR[NBUF] = {}; float* d_dep[NBUF] = {}; int* d_mi[NBUF] = {};
    bool host_ready = false;
    long long host_bytes = 0;
    int* h_status = nullptr; int h_status_cap = 0, h_status_n = 0;   // pinned: per-pair range-error flags of the last run_host
    int copy_only = 0;        // diagnostic: run_host moves its bytes but launches no kernel (host-link ceiling of the same traffic pattern)
    int status_bank = 0;      // which half of eng.d_status the next run_device uses (run_host: the chunk's slot, so that the
                              // previous chunk's flag download on the output stream never races with the next chunk's clear)
    int stereo_flags = 0;
    int* d_sad[2] = {nullptr, nullptr};     // [P][C] per compute lane, only with the median cull
    // per-kernel timing (b200orb_batch_profile): a pool of event sets, one set per run_device call
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_pairs;
    int prof_cap = 0, prof_used = 0;
    bool prof_on = false;
};

extern "C" {

const char* b200orb_last_error(void) { return g_err.c_str(); }
int b200orb_version(void) { return 100; }
int b200orb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
long long b200orb_kernel_launches(void) { return g_launches.load(); }

int b200orb_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device,
                             b200orb_extractor** out) {
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    *out = nullptr;
    b200orb_extractor* e = new b200orb_extractor;
    int r = make_params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, e->eng.prm);
    if (r) { delete e; return r; }
    e->eng.device = device;
    *out = e;
    return 0;
}

void b200orb_extractor_destroy(b200orb_extractor* e) {
    if (!e) return;
    if (e->st || e->d_img || e->eng.planned) {
        cudaSetDevice(e->eng.device);
        if (e->st) cudaStreamSynchronize(e->st);
        if (e->st_pyr) cudaStreamSynchronize(e->st_pyr);
        e->drop_graph();
        e->eng.release();
        cudaFree(e->d_img); cudaFree(e->d_out); cudaFree(e->d_sout);
        if (e->h_out) cudaFreeHost(e->h_out);
        if (e->h_sout) cudaFreeHost(e->h_sout);
        if (e->h_img) cudaFreeHost(e->h_img);
        if (e->h_pyr) cudaFreeHost(e->h_pyr);
        if (e->ev_done) cudaEventDestroy(e->ev_done);
        if (e->ev_pyr) cudaEventDestroy(e->ev_pyr);
        if (e->st) cudaStreamDestroy(e->st);
        if (e->st_pyr) cudaStreamDestroy(e->st_pyr);
    }
    delete e;
}

int b200orb_get_levels(const b200orb_extractor* e) { return e->eng.prm.nlevels; }
float b200orb_get_scale_factor(const b200orb_extractor* e) { return e->eng.prm.scaleFactorF; }
static int copy_tab(const std::vector<float>& v, float* out) { if (!out) return fail(B200ORB_E_ARG, "out is NULL"); memcpy(out, v.data(), v.size() * 4); return 0; }
int b200orb_get_scale_factors(const b200orb_extractor* e, float* out) { return copy_tab(e->eng.prm.sf, out); }
int b200orb_get_inverse_scale_factors(const b200orb_extractor* e, float* out) { return copy_tab(e->eng.prm.isf, out); }
int b200orb_get_scale_sigma_squares(const b200orb_extractor* e, float* out) { return copy_tab(e->eng.prm.sig2, out); }
int b200orb_get_inverse_scale_sigma_squares(const b200orb_extractor* e, float* out) { return copy_tab(e->eng.prm.isig2, out); }
int b200orb_get_features_per_level(const b200orb_extractor* e, int* out) {
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    memcpy(out, e->eng.prm.quota.data(), e->eng.prm.quota.size() * 4);
    return 0;
}

static size_t out_desc_ofs(int C) { return 256 + (((size_t)C * 24 + 255) & ~(size_t)255); }     // descriptors are read as 16-byte vectors
static size_t out_block_bytes(int C) { return out_desc_ofs(C) + (size_t)C * 32; }

int b200orb_extract(b200orb_extractor* e, const uint8_t* image, int H, int W, int* n_keypoints) {
    if (!e || !n_keypoints) return fail(B200ORB_E_ARG, "NULL argument");
    if (H == 0 || W == 0 || !image) {       // ORBextractor.cpp:1045-1046: empty image -> nothing happens
        *n_keypoints = 0; e->n = 0; e->empty_last = true;
        return 0;
    }
    if (H < 0 || W < 0) return fail(B200ORB_E_ARG, "negative image size");
    CU_TRY(cudaSetDevice(e->eng.device));
    if (!e->st) {
        CU_TRY(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&e->st_pyr, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&e->ev_pyr, cudaEventDisableTiming));
    }
    if (e->use_graph < 0) { const char* v = getenv("B200ORB_GRAPH"); e->use_graph = v ? atoi(v) : 1; }
    if (e->pyr_inflight) {                  // the previous call's pyramid download reads the blob the new call overwrites
        CU_TRY(cudaStreamSynchronize(e->st_pyr));
        e->pyr_inflight = false;
    }
    e->pyr_valid = false;
    const bool replanned = !(e->eng.planned && e->eng.hp.P.H == H && e->eng.hp.P.W == W);
    TRY(e->eng.plan(H, W, 1));
    const Plan& P = e->eng.hp.P;
    const size_t HW = (size_t)H * W;
    if (replanned || HW > e->img_cap || e->out_cap != P.kp_total) e->drop_graph();     // the graph holds the old buffers' addresses
    if (HW > e->img_cap) {
        cudaFree(e->d_img); e->d_img = nullptr;
        if (e->h_img) cudaFreeHost(e->h_img);
        e->h_img = nullptr;
        CU_TRY(cudaMalloc((void**)&e->d_img, HW));
        CU_TRY(cudaHostAlloc((void**)&e->h_img, HW, cudaHostAllocDefault));
        e->img_cap = HW; e->h_img_cap = HW;
    }
    if (e->out_cap != P.kp_total) {
        const int C = P.kp_total;
        cudaFree(e->d_out); cudaFree(e->d_sout);
        if (e->h_out) cudaFreeHost(e->h_out);
        if (e->h_sout) cudaFreeHost(e->h_sout);
        e->d_out = e->d_sout = nullptr; e->h_out = e->h_sout = nullptr; e->out_cap = 0;
        CU_TRY(cudaMalloc((void**)&e->d_out, out_block_bytes(C)));
        CU_TRY(cudaHostAlloc((void**)&e->h_out, out_block_bytes(C), cudaHostAllocDefault));
        CU_TRY(cudaMalloc((void**)&e->d_sout, (size_t)C * 16 + 16));
        CU_TRY(cudaHostAlloc((void**)&e->h_sout, (size_t)C * 16 + 16, cudaHostAllocDefault));
        e->d_nkp = (int*)e->d_out; e->d_kps = (float*)(e->d_out + 256); e->d_desc = e->d_out + out_desc_ofs(C);
        e->d_uR = (float*)e->d_sout; e->d_depth = e->d_uR + C; e->d_match = (int*)(e->d_depth + C); e->d_sad = e->d_match + C;
        e->d_sstatus = e->d_sad + C;
        e->out_cap = C;
    }
    if ((size_t)P.pyr_bytes > e->h_pyr_cap) {
        if (e->h_pyr) cudaFreeHost(e->h_pyr);
        e->h_pyr = nullptr; e->h_pyr_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&e->h_pyr, (size_t)P.pyr_bytes, cudaHostAllocDefault));
        e->h_pyr_cap = (size_t)P.pyr_bytes;
    }
    memcpy(e->h_img, image, HW);            // the caller's array is pageable; one host copy into the pinned staging buffer
    auto enqueue = [&]() -> int {
        CU_TRY(cudaMemcpyAsync(e->d_img, e->h_img, HW, cudaMemcpyHostToDevice, e->st));
        TRY(e->eng.extract(e->d_img, e->d_img, 1, 1, e->d_kps, e->d_desc, e->d_nkp, e->st));
        CU_TRY(cudaMemcpyAsync(e->h_out, e->d_out, out_block_bytes(P.kp_total), cudaMemcpyDeviceToHost, e->st));
        return 0;
    };
    if (e->use_graph) {
        if (!e->graph) {
            cudaGraph_t g = nullptr;
            const long long before = g_launches.load();
            CU_TRY(cudaStreamBeginCapture(e->st, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue();
            e->graph_kernels = g_launches.load() - before;
            g_launches -= e->graph_kernels;          // counted per launch of the graph below
            const cudaError_t ce = cudaStreamEndCapture(e->st, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) return fail(B200ORB_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&e->graph, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { e->graph = nullptr; return fail(B200ORB_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
            e->graph_H = H; e->graph_W = W;
        }
        CU_TRY(cudaGraphLaunch(e->graph, e->st));
        g_launches += e->graph_kernels;
    } else {
        TRY(enqueue());
    }
    CU_TRY(cudaEventRecord(e->ev_done, e->st));
    if (e->prefetch_pyramid) {              // the pyramid follows on its own stream while the caller already works on the keypoints
        CU_TRY(cudaStreamWaitEvent(e->st_pyr, e->ev_done, 0));
        CU_TRY(cudaMemcpyAsync(e->h_pyr, e->eng.d_pyr, (size_t)P.pyr_bytes, cudaMemcpyDeviceToHost, e->st_pyr));
        e->pyr_inflight = true;
    }
    CU_TRY(cudaEventSynchronize(e->ev_done));
    e->n = *reinterpret_cast<const int*>(e->h_out);
    e->empty_last = false;
    *n_keypoints = e->n;
    return 0;
}

int b200orb_get_results(b200orb_extractor* e, float* kps, uint8_t* desc) {
    if (!e || e->n < 0) return fail(B200ORB_E_STATE, "no extract() call yet");
    if (e->n == 0) return 0;
    // the results already sit in the pinned mirror (one copy at the end of the launch sequence)
    if (kps) memcpy(kps, e->h_out + 256, (size_t)e->n * 24);
    if (desc) memcpy(desc, e->h_out + out_desc_ofs(e->out_cap), (size_t)e->n * 32);
    return 0;
}

int b200orb_max_keypoints(const b200orb_extractor* e) { return e && e->eng.planned ? e->eng.hp.P.kp_total : 0; }

static int need_pyramid(const b200orb_extractor* e, int level) {
    if (!e || e->n < 0 || e->empty_last || !e->eng.planned) return fail(B200ORB_E_STATE, "no pyramid: extract() has not processed an image yet");
    if (level < 0 || level >= e->eng.prm.nlevels) return fail(B200ORB_E_ARG, "level out of range");
    return 0;
}

int b200orb_level_size(const b200orb_extractor* e, int level, int* w, int* h) {
    TRY(need_pyramid(e, level));
    *w = e->eng.hp.P.lv[level].w; *h = e->eng.hp.P.lv[level].h;
    return 0;
}

// host mirror of the pyramid blob (physical layout) of the last extract; downloaded once, in the background when prefetch is on
static int fetch_pyramid(b200orb_extractor* e) {
    if (e->pyr_valid) return 0;
    CU_TRY(cudaSetDevice(e->eng.device));
    if (!e->pyr_inflight) {
        CU_TRY(cudaMemcpyAsync(e->h_pyr, e->eng.d_pyr, (size_t)e->eng.hp.P.pyr_bytes, cudaMemcpyDeviceToHost, e->st_pyr));
    }
    CU_TRY(cudaStreamSynchronize(e->st_pyr));
    e->pyr_inflight = false;
    e->pyr_valid = true;
    return 0;
}
// the caster's step-ignoring copy (opencv_type_casters.h:230-239): rows*cols contiguous LOGICAL bytes from the ROI start of the
// (w + 38)-pitch bordered buffer; the mirror has the padded physical pitch, so the run is gathered row by row
static void sheared_view(const b200orb_extractor* e, int level, uint8_t* out) {
    const LevelGeom& G = e->eng.hp.P.lv[level];
    const int plog = G.w + 2 * ORB_EDGE;
    const u8* base = e->h_pyr + G.pyr_ofs;
    size_t lin = (size_t)ORB_EDGE * plog + ORB_EDGE, left = (size_t)G.w * G.h;
    while (left) {
        const size_t pr = lin / plog, pc = lin - pr * plog, n = std::min(left, (size_t)plog - pc);
        memcpy(out, base + pr * G.pitch + pc, n);
        out += n; lin += n; left -= n;
    }
}

int b200orb_get_pyramid_level(b200orb_extractor* e, int level, uint8_t* out) {
    TRY(need_pyramid(e, level));
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    TRY(fetch_pyramid(e));
    sheared_view(e, level, out);
    return 0;
}

int b200orb_get_pyramid_all(b200orb_extractor* e, uint8_t* out, long long cap) {
    TRY(need_pyramid(e, 0));
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    const Plan& P = e->eng.hp.P;
    size_t total = 0;
    for (int l = 0; l < P.nlevels; ++l) total += (size_t)P.lv[l].w * P.lv[l].h;
    if ((long long)total > cap) return fail(B200ORB_E_ARG, "output buffer too small for the pyramid views");
    TRY(fetch_pyramid(e));
    size_t o = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        sheared_view(e, l, out + o);
        o += (size_t)P.lv[l].w * P.lv[l].h;
    }
    return 0;
}

int b200orb_get_level_image(b200orb_extractor* e, int level, int blurred, uint8_t* out) {
    TRY(need_pyramid(e, level));
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    CU_TRY(cudaSetDevice(e->eng.device));
    const LevelGeom& G = e->eng.hp.P.lv[level];
    if (blurred)
        CU_TRY(cudaMemcpy2DAsync(out, G.w, e->eng.d_blur + G.blur_ofs, G.blur_pitch, G.w, G.h, cudaMemcpyDeviceToHost, e->st));
    else
        CU_TRY(cudaMemcpy2DAsync(out, G.w, e->eng.d_pyr + G.pyr_ofs + (size_t)ORB_EDGE * G.pitch + ORB_EDGE, G.pitch, G.w, G.h,
                                 cudaMemcpyDeviceToHost, e->st));
    CU_TRY(cudaStreamSynchronize(e->st));
    return 0;
}

int b200orb_get_level_candidates(b200orb_extractor* e, int level, int cap, int* out, int* n) {
    TRY(need_pyramid(e, level));
    if (!n) return fail(B200ORB_E_ARG, "n is NULL");
    CU_TRY(cudaSetDevice(e->eng.device));
    const Plan& P = e->eng.hp.P;
    const LevelGeom& G = P.lv[level];
    const int nc = G.nRows * G.nCols;
    std::vector<int> cnt(std::max(nc, 1));
    std::vector<u32> c((size_t)std::max(nc * G.cell_cap, 1));
    if (nc) {
        CU_TRY(cudaMemcpy(cnt.data(), e->eng.d_cellcnt + G.cell_ofs, (size_t)nc * 4, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(c.data(), e->eng.d_cand + G.cand_ofs, (size_t)nc * G.cell_cap * 4, cudaMemcpyDeviceToHost));
    }
    int k = 0;
    for (int i = 0; i < nc; ++i)
        for (int j = 0; j < cnt[i]; ++j, ++k)
            if (out && k < cap) {
                const u32 v = c[(size_t)i * G.cell_cap + j];
                out[3 * k] = v & 0xfff; out[3 * k + 1] = (v >> 12) & 0xfff; out[3 * k + 2] = v >> 24;
            }
    *n = k;
    return 0;
}

int b200orb_stereo(b200orb_extractor* L, b200orb_extractor* R, double mbf, float fx, float* uRight, float* depth, int* matchIdx) {
    return b200orb_stereo_ex(L, R, mbf, fx, 0, uRight, depth, matchIdx, nullptr);
}

int b200orb_stereo_ex(b200orb_extractor* L, b200orb_extractor* R, double mbf, float fx, int flags, float* uRight, float* depth, int* matchIdx,
                      int* sadDist) {
    if (!L || !R || !uRight || !depth) return fail(B200ORB_E_ARG, "NULL argument");
    if (L->n < 0 || R->n < 0) return fail(B200ORB_E_STATE, "both extractors need an extract() call first");
    if (L->n == 0) return 0;
    if (L->empty_last || R->empty_last) {
        if (R->empty_last && !L->empty_last) { for (int i = 0; i < L->n; ++i) { uRight[i] = -1.f; depth[i] = -1.f; if (matchIdx) matchIdx[i] = -1; } return 0; }
        return fail(B200ORB_E_STATE, "left extractor holds no image");
    }
    if (L->eng.device != R->eng.device) return fail(B200ORB_E_ARG, "extractors live on different devices");
    const Plan &PL = L->eng.hp.P, &PR = R->eng.hp.P;
    if (PL.H != PR.H || PL.W != PR.W || PL.nlevels != PR.nlevels || L->eng.prm.scaleFactorF != R->eng.prm.scaleFactorF)
        return fail(B200ORB_E_ARG, "left and right extractors have different geometry");
    CU_TRY(cudaSetDevice(L->eng.device));
    StereoGeom SG;
    L->eng.stereo_geom(SG, flags);
    StereoArgs A;
    memset(&A, 0, sizeof(A));
    A.kpsL = L->d_kps; A.descL = L->d_desc; A.nL = L->d_nkp;
    A.kpsR = R->d_kps; A.descR = R->d_desc; A.nR = R->d_nkp;
    A.pyrL = L->eng.d_pyr; A.pyrR = R->eng.d_pyr;
    A.kp_row = 6; A.oct_idx = 5; A.out_stride = PL.kp_total;
    A.uRight = L->d_uR; A.depth = L->d_depth; A.matchIdx = L->d_match; A.status = L->d_sstatus; A.sadDist = L->d_sad;
    A.rowStart = L->eng.d_rowstart; A.rmeta = L->eng.d_rmeta; A.idx_stride = (long long)PL.kp_total * L->eng.band_rows();
    fill_stereo_consts(A, mbf, fx);
    const size_t C = (size_t)PL.kp_total;
    CU_TRY(cudaMemsetAsync(L->d_sstatus, 0, 4, L->st));
    TRY(launch_stereo(SG, A, L->n, 1, L->st, flags));
    CU_TRY(cudaMemcpyAsync(L->h_sout, L->d_sout, C * 16 + 16, cudaMemcpyDeviceToHost, L->st));     // all four arrays + the status word
    CU_TRY(cudaStreamSynchronize(L->st));
    const size_t nb = (size_t)L->n * 4;
    memcpy(uRight, L->h_sout, nb);
    memcpy(depth, L->h_sout + C * 4, nb);
    if (matchIdx) memcpy(matchIdx, L->h_sout + C * 8, nb);
    if (sadDist) memcpy(sadDist, L->h_sout + C * 12, nb);
    const int status = *reinterpret_cast<const int*>(L->h_sout + C * 16);
    if (status) return fail(B200ORB_E_RANGE, "a SAD window leaves the pyramid view (the reference raises IndexError/ValueError here)");
    return 0;
}

int b200orb_stereo_host(int device, int nLeft, const float* kpsL, const uint8_t* descL, int nRight, const float* kpsR,
                        const uint8_t* descR, int nlevels, const float* sf, const float* isf, const uint8_t* const* pyrL,
                        const uint8_t* const* pyrR, const int* lw, const int* lh, double mbf, float fx, float* uRight, float* depth,
                        int* matchIdx) {
    if (nLeft < 0 || nRight < 0 || nlevels < 1 || nlevels > ORB_MAX_LEVELS) return fail(B200ORB_E_ARG, "bad sizes");
    if (nLeft == 0) return 0;
    if (nRight >= (1 << 20)) return fail(B200ORB_E_ARG, "more than 2^20 right keypoints");
    if (lh[0] < 1 || lh[0] > 4096) return fail(B200ORB_E_ARG, "level-0 height must be in [1, 4096]");
    CU_TRY(cudaSetDevice(device));
    StereoGeom SG;
    memset(&SG, 0, sizeof(SG));
    SG.nlevels = nlevels; SG.nRows = lh[0];
    long long total = 0;
    for (int l = 0; l < nlevels; ++l) {
        SG.sf[l] = sf[l]; SG.isf[l] = isf[l]; SG.w[l] = lw[l]; SG.h[l] = lh[l];
        SG.plog[l] = lw[l]; SG.pitch[l] = lw[l]; SG.off0[l] = 0; SG.base[l] = total; SG.vstride[l] = lw[l];
        SG.magic[l] = lw[l] > 1 ? (unsigned)((0x100000000ULL + lw[l] - 1) / lw[l]) : 0xffffffffu;
        total += ((long long)lw[l] * lh[l] + 255) & ~255LL;
    }
    // host-side validation of what the reference would index (rows of vRowIndices, octaves)
    for (int i = 0; i < nRight; ++i) {
        const int o = (int)kpsR[3 * i + 2];
        if (o < 0 || o >= nlevels) return fail(B200ORB_E_RANGE, "right keypoint octave out of range");
        const double y = kpsR[3 * i + 1], r = 2.0 * (double)sf[o];
        if (floor(y - r) < 0 || ceil(y + r) >= lh[0]) return fail(B200ORB_E_RANGE, "right keypoint row band leaves the image (the reference raises IndexError)");
    }
    for (int i = 0; i < nLeft; ++i) {
        const int o = (int)kpsL[3 * i + 2];
        if (o < 0 || o >= nlevels) return fail(B200ORB_E_RANGE, "left keypoint octave out of range");
        if ((int)kpsL[3 * i + 1] < 0 || (int)kpsL[3 * i + 1] >= lh[0]) return fail(B200ORB_E_RANGE, "left keypoint row outside the image");
    }
    // Device scratch of this entry point is cached per device and only grows (the first version paid eleven cudaMalloc / cudaFree
    // pairs per call -- ~10 ms before any work).  One arena, carved into 256-byte aligned pieces; the uploads are plain copies of the
    // caller's (pageable) arrays.
    struct HostStereoWS { std::mutex mu; u8* base = nullptr; size_t cap = 0; cudaStream_t st = nullptr; };
    static std::mutex ws_mutex;
    static std::map<int, HostStereoWS> ws_map;
    HostStereoWS* wsp;
    { std::lock_guard<std::mutex> g(ws_mutex); wsp = &ws_map[device]; }      // std::map nodes are stable
    HostStereoWS& ws = *wsp;
    std::lock_guard<std::mutex> guard(ws.mu);                                 // calls on one device are serialised, devices run side by side
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t nR1 = (size_t)std::max(nRight, 1);
    float smax_h = 1.f;
    for (int l = 0; l < nlevels; ++l) smax_h = std::max(smax_h, sf[l]);
    const size_t band = (size_t)(2 * (int)ceil(2.0 * smax_h) + 2);
    const size_t o_blob = 0, o_kL = o_blob + al((size_t)total * 2), o_dL = o_kL + al((size_t)nLeft * 12), o_kR = o_dL + al((size_t)nLeft * 32),
                 o_dR = o_kR + al(nR1 * 12), o_out = o_dR + al(nR1 * 32), o_n = o_out + al((size_t)nLeft * 12), o_rs = o_n + 256,
                 o_rm = o_rs + al((size_t)(lh[0] + 1) * 4), need = o_rm + al(nR1 * band * 8);
    if (need > ws.cap) {
        if (ws.base) cudaFree(ws.base);
        ws.base = nullptr; ws.cap = 0;
        CU_TRY(cudaMalloc((void**)&ws.base, need + need / 4));
        ws.cap = need + need / 4;
    }
    if (!ws.st) CU_TRY(cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking));
    cudaStream_t st = ws.st;
    u8* d_blob = ws.base + o_blob;
    float* d_kL = (float*)(ws.base + o_kL); u8* d_dL = ws.base + o_dL;
    float* d_kR = (float*)(ws.base + o_kR); u8* d_dR = ws.base + o_dR;
    float* d_u = (float*)(ws.base + o_out); float* d_d = d_u + nLeft; int* d_m = (int*)(d_d + nLeft);
    int* d_n = (int*)(ws.base + o_n); int* d_rs = (int*)(ws.base + o_rs); uint2* d_rm = (uint2*)(ws.base + o_rm);
    for (int l = 0; l < nlevels; ++l) {
        CU_TRY(cudaMemcpyAsync(d_blob + SG.base[l], pyrL[l], (size_t)lw[l] * lh[l], cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_blob + total + SG.base[l], pyrR[l], (size_t)lw[l] * lh[l], cudaMemcpyHostToDevice, st));
    }
    CU_TRY(cudaMemcpyAsync(d_kL, kpsL, (size_t)nLeft * 12, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_dL, descL, (size_t)nLeft * 32, cudaMemcpyHostToDevice, st));
    if (nRight) {
        CU_TRY(cudaMemcpyAsync(d_kR, kpsR, (size_t)nRight * 12, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_dR, descR, (size_t)nRight * 32, cudaMemcpyHostToDevice, st));
    }
    const int hn[3] = {nLeft, nRight, 0};
    CU_TRY(cudaMemcpyAsync(d_n, hn, 12, cudaMemcpyHostToDevice, st));
    StereoArgs A;
    memset(&A, 0, sizeof(A));
    A.kpsL = d_kL; A.descL = d_dL; A.nL = d_n; A.kpsR = d_kR; A.descR = d_dR; A.nR = d_n + 1;
    A.pyrL = d_blob; A.pyrR = d_blob + total;
    A.kp_row = 3; A.oct_idx = 2; A.out_stride = nLeft;
    A.uRight = d_u; A.depth = d_d; A.matchIdx = d_m; A.status = d_n + 2;
    A.rowStart = d_rs; A.rmeta = d_rm; A.idx_stride = (long long)(nR1 * band);
    fill_stereo_consts(A, mbf, fx);
    TRY(launch_stereo(SG, A, nLeft, 1, st));
    int status = 0;
    CU_TRY(cudaMemcpyAsync(uRight, d_u, (size_t)nLeft * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(depth, d_d, (size_t)nLeft * 4, cudaMemcpyDeviceToHost, st));
    if (matchIdx) CU_TRY(cudaMemcpyAsync(matchIdx, d_m, (size_t)nLeft * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&status, d_n + 2, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (status) return fail(B200ORB_E_RANGE, "a SAD window leaves the pyramid view (the reference raises IndexError/ValueError here)");
    return 0;
}

// ---------------------------------------------------------------- batch
int b200orb_batch_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int H, int W, int max_pairs,
                         int device, b200orb_batch** out) {
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    *out = nullptr;
    if (max_pairs < 1) return fail(B200ORB_E_ARG, "max_pairs must be >= 1");
    b200orb_batch* b = new b200orb_batch;
    int r = make_params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, b->eng.prm);
    if (!r) { b->eng.device = device; b->P = max_pairs; b->H = H; b->W = W; r = b->eng.plan(H, W, 2 * max_pairs); }
    if (r) { b200orb_batch_destroy(b); return r; }
    *out = b;
    return 0;
}

void b200orb_batch_destroy(b200orb_batch* b) {
    if (!b) return;
    cudaSetDevice(b->eng.device);
    b->eng.release();
    if (b->eng2) { b->eng2->release(); delete b->eng2; }
    if (b->s_comp2) cudaStreamDestroy(b->s_comp2);
    for (cudaEvent_t e : b->prof_ev) cudaEventDestroy(e);
    cudaFree(b->d_sad[0]); cudaFree(b->d_sad[1]);
    for (int i = 0; i < b200orb_batch::NBUF; ++i) {
        cudaFree(b->d_in[i]); cudaFree(b->d_kps[i]); cudaFree(b->d_desc[i]); cudaFree(b->d_nkp[i]);
        cudaFree(b->d_uR[i]); cudaFree(b->d_dep[i]); cudaFree(b->d_mi[i]);
        if (b->ev_in[i]) cudaEventDestroy(b->ev_in[i]);
        if (b->ev_comp[i]) cudaEventDestroy(b->ev_comp[i]);
        if (b->ev_out[i]) cudaEventDestroy(b->ev_out[i]);
    }
    if (b->h_status) cudaFreeHost(b->h_status);
    if (b->s_in) cudaStreamDestroy(b->s_in);
    if (b->s_comp) cudaStreamDestroy(b->s_comp);
    if (b->s_out) cudaStreamDestroy(b->s_out);
    delete b;
}

int b200orb_batch_max_pairs(const b200orb_batch* b) { return b ? b->P : 0; }
int b200orb_batch_kp_capacity(const b200orb_batch* b) { return b ? b->eng.hp.P.kp_total : 0; }
long long b200orb_batch_workspace_bytes(const b200orb_batch* b) { return b ? b->eng.bytes + (b->eng2 ? b->eng2->bytes : 0) + b->host_bytes : 0; }

static int batch_run_on(b200orb_batch* b, Engine& E, const uint8_t* d_left, const uint8_t* d_right, int n_pairs, double mbf, float fx,
                        float* d_kps, uint8_t* d_desc, int32_t* d_nkp, float* d_uRight, float* d_depth, int32_t* d_matchIdx,
                        void* stream) {
    if (!b || !d_left || !d_right || !d_kps || !d_desc || !d_nkp || !d_uRight || !d_depth) return fail(B200ORB_E_ARG, "NULL argument");
    if (n_pairs < 1 || n_pairs > b->P) return fail(B200ORB_E_ARG, "n_pairs must be in [1, max_pairs]");
    CU_TRY(cudaSetDevice(E.device));
    cudaStream_t st = (cudaStream_t)stream;
    const Plan& P = E.hp.P;
    const size_t C = P.kp_total;
    cudaEvent_t* evs = nullptr;
    if (b->prof_on && b->prof_used < b->prof_cap) {
        evs = b->prof_ev.data() + (size_t)b->prof_used * (B200ORB_NSTAGE + 1);
        b->prof_pairs[b->prof_used] = n_pairs;
        ++b->prof_used;
        CU_TRY(cudaEventRecord(evs[0], st));
    }
    TRY(E.extract(d_left, d_right, n_pairs, 2 * n_pairs, d_kps, d_desc, d_nkp, st, evs));
    StereoGeom SG;
    E.stereo_geom(SG, b->stereo_flags);
    StereoArgs A;
    memset(&A, 0, sizeof(A));
    A.kpsL = d_kps; A.kpsR = d_kps + (size_t)n_pairs * C * 6;
    A.descL = d_desc; A.descR = d_desc + (size_t)n_pairs * C * 32;
    A.nL = d_nkp; A.nR = d_nkp + n_pairs; A.n_stride = 1;
    A.pyrL = E.d_pyr; A.pyrR = E.d_pyr + (size_t)n_pairs * P.pyr_bytes;
    A.kp_stride = (long long)C * 6; A.desc_stride = (long long)C * 32; A.pyr_stride = P.pyr_bytes;
    A.kp_row = 6; A.oct_idx = 5; A.out_stride = (int)C;
    A.uRight = d_uRight; A.depth = d_depth; A.matchIdx = d_matchIdx;
    A.status = E.d_status + (size_t)b->status_bank * E.S; A.status_stride = 1;   // flag word per pair, cleared by k_rowindex, read by b200orb_batch_status_device / run_host
    A.rowStart = E.d_rowstart; A.rmeta = E.d_rmeta; A.idx_stride = (long long)C * E.band_rows();
    fill_stereo_consts(A, mbf, fx);
    if (b->stereo_flags & B200ORB_STEREO_MEDIAN_CULL) {
        int*& sad = b->d_sad[&E == b->eng2 ? 1 : 0];
        if (!sad) CU_TRY(cudaMalloc((void**)&sad, (size_t)b->P * C * 4));
        A.sadDist = sad;
    }
    TRY(launch_stereo(SG, A, (int)C, n_pairs, st, b->stereo_flags));
    if (evs) CU_TRY(cudaEventRecord(evs[B200ORB_NSTAGE], st));
    return 0;
}

int b200orb_batch_run_device(b200orb_batch* b, const uint8_t* d_left, const uint8_t* d_right, int n_pairs, double mbf, float fx,
                             float* d_kps, uint8_t* d_desc, int32_t* d_nkp, float* d_uRight, float* d_depth, int32_t* d_matchIdx,
                             void* stream) {
    if (!b) return fail(B200ORB_E_ARG, "NULL argument");
    return batch_run_on(b, b->eng, d_left, d_right, n_pairs, mbf, fx, d_kps, d_desc, d_nkp, d_uRight, d_depth, d_matchIdx, stream);
}

static const char* kRangeMsg = "a SAD window or row band leaves the pyramid view in at least one pair (the reference raises IndexError/ValueError "
                               "there, Frame.py:230-250); the per-pair flags say which";

int b200orb_batch_status_device(b200orb_batch* b, int n_pairs, void* stream, int32_t* pair_status) {
    if (!b) return fail(B200ORB_E_ARG, "NULL batch");
    if (n_pairs < 1 || n_pairs > b->P) return fail(B200ORB_E_ARG, "n_pairs must be in [1, max_pairs]");
    CU_TRY(cudaSetDevice(b->eng.device));
    std::vector<int> tmp(n_pairs);
    CU_TRY(cudaMemcpyAsync(tmp.data(), b->eng.d_status, (size_t)n_pairs * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    int any = 0;
    for (int i = 0; i < n_pairs; ++i) { any |= tmp[i]; if (pair_status) pair_status[i] = tmp[i]; }
    return any ? fail(B200ORB_E_RANGE, kRangeMsg) : 0;
}

int b200orb_batch_status_host(const b200orb_batch* b, int32_t* pair_status, int n_pairs) {
    if (!b || !pair_status) return fail(B200ORB_E_ARG, "NULL argument");
    if (n_pairs < 0 || n_pairs > b->h_status_n) return fail(B200ORB_E_ARG, "n_pairs exceeds the last run_host job");
    memcpy(pair_status, b->h_status, (size_t)n_pairs * sizeof(int));
    return 0;
}

int b200orb_batch_set_copy_only(b200orb_batch* b, int on) {
    if (!b) return fail(B200ORB_E_ARG, "NULL batch");
    b->copy_only = on ? 1 : 0;
    return 0;
}

int b200orb_batch_set_stereo_flags(b200orb_batch* b, int flags) {
    if (!b) return fail(B200ORB_E_ARG, "NULL batch");
    if (flags & ~(B200ORB_STEREO_MEDIAN_CULL | B200ORB_STEREO_DENSE_PYRAMID)) return fail(B200ORB_E_ARG, "unknown stereo flag");
    b->stereo_flags = flags;
    return 0;
}

int b200orb_batch_candidate_count(b200orb_batch* b, int n_images, long long* total) {
    if (!b || !total) return fail(B200ORB_E_ARG, "NULL argument");
    if (n_images < 1 || n_images > b->eng.S) return fail(B200ORB_E_ARG, "n_images out of range");
    CU_TRY(cudaSetDevice(b->eng.device));
    const Plan& P = b->eng.hp.P;
    std::vector<int> cnt((size_t)n_images * P.ncells);
    CU_TRY(cudaMemcpy(cnt.data(), b->eng.d_cellcnt, cnt.size() * 4, cudaMemcpyDeviceToHost));
    long long t = 0;
    int real_cells = 0;
    for (int l = 0; l < P.nlevels; ++l) real_cells += P.lv[l].nRows * P.lv[l].nCols;
    for (int i = 0; i < n_images; ++i)
        for (int c = 0; c < real_cells; ++c) t += cnt[(size_t)i * P.ncells + c];
    *total = t;
    return 0;
}

int b200orb_batch_profile(b200orb_batch* b, int enable, int max_calls) {
    if (!b) return fail(B200ORB_E_ARG, "NULL batch");
    CU_TRY(cudaSetDevice(b->eng.device));
    if (enable && max_calls > b->prof_cap) {
        const size_t want = (size_t)max_calls * (B200ORB_NSTAGE + 1);
        while (b->prof_ev.size() < want) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreate(&e));
            b->prof_ev.push_back(e);
        }
        b->prof_pairs.resize(max_calls);
        b->prof_cap = max_calls;
    }
    b->prof_on = enable != 0;
    b->prof_used = 0;
    return 0;
}

int b200orb_batch_stage_launches(const b200orb_batch* b, int* launches_per_stage) {
    if (!b || !launches_per_stage) return fail(B200ORB_E_ARG, "NULL argument");
    const int v[B200ORB_NSTAGE] = {1, b->eng.hp.P.nlevels - 1, 1, b->eng.hp.fast_cells > 0 ? 1 : 0, 1, 1, 2 + ((b->stereo_flags & B200ORB_STEREO_MEDIAN_CULL) ? 1 : 0)};
    for (int k = 0; k < B200ORB_NSTAGE; ++k) launches_per_stage[k] = v[k];
    return 0;
}

int b200orb_batch_profile_read(b200orb_batch* b, float* ms_per_stage, int* n_calls, long long* n_pairs) {
    if (!b || !ms_per_stage || !n_calls) return fail(B200ORB_E_ARG, "NULL argument");
    CU_TRY(cudaSetDevice(b->eng.device));
    for (int k = 0; k < B200ORB_NSTAGE; ++k) ms_per_stage[k] = 0.f;
    long long pairs = 0;
    for (int c = 0; c < b->prof_used; ++c) {
        cudaEvent_t* evs = b->prof_ev.data() + (size_t)c * (B200ORB_NSTAGE + 1);
        CU_TRY(cudaEventSynchronize(evs[B200ORB_NSTAGE]));
        for (int k = 0; k < B200ORB_NSTAGE; ++k) {
            float ms = 0.f;
            CU_TRY(cudaEventElapsedTime(&ms, evs[k], evs[k + 1]));
            ms_per_stage[k] += ms;
        }
        pairs += b->prof_pairs[c];
    }
    *n_calls = b->prof_used;
    if (n_pairs) *n_pairs = pairs;
    b->prof_used = 0;
    return 0;
}

// Chunk schedule of run_host: full chunks with a short ramp at both ends of a long job (C/4, C/2, C ... C, C/2, C/4) so that the
// first kernels start after a quarter-chunk upload and only a quarter-chunk download is left when the last kernels finish.
// With two compute lanes the chunk C is half the engine's capacity: a chunk's kernels then take longer than the next chunk's
// upload, so consecutive chunks do overlap on the GPU (with full chunks the next upload ends just as the kernels do), and the
// job ends half as long after its last upload.
static std::vector<int> host_chunk_schedule(int max_pairs, int lanes, int n_pairs) {
    std::vector<int> sizes, tail;
    const int P = (lanes == 2 && max_pairs >= 16) ? max_pairs / 2 : max_pairs;
    int left = n_pairs;
    if (n_pairs >= 4 * P && P >= 8) {
        static const int min_tail = [] { const char* v = getenv("B200ORB_HOST_MIN_TAIL"); return v ? std::max(1, atoi(v)) : 0; }();
        sizes.push_back(P / 4); sizes.push_back(P / 2);
        left -= P / 4 + P / 2;
        // the job ends one chunk latency after its last upload: taper the last chunks down to min_tail pairs (default C/4)
        for (int c = P / 2; c >= std::max(min_tail ? min_tail : P / 4, 1) && c >= 4; c /= 2) { tail.push_back(c); left -= c; }
    }
    while (left > 0) { const int c = std::min(P, left); sizes.push_back(c); left -= c; }
    sizes.insert(sizes.end(), tail.begin(), tail.end());
    return sizes;
}

int b200orb_plan_cells(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int H, int W, int32_t* cells,
                       int capacity) {
    Params prm;
    TRY(make_params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, prm));
    HostPlan* hp = new HostPlan;
    const int r = build_plan(prm, H, W, *hp);
    if (r) { delete hp; return r; }
    const int n = hp->fast_cells;
    for (int i = 0; cells && i < n && i < capacity; ++i) {
        const uint4 t = hp->celltab[i];
        int32_t* o = cells + 6 * (size_t)i;
        o[0] = (int)(t.y >> 16); o[1] = (int)(t.x & 0xffffu); o[2] = (int)(t.x >> 16);
        o[3] = (int)(t.y & 0xffu); o[4] = (int)((t.y >> 8) & 0xffu); o[5] = (int)t.z;
    }
    delete hp;
    return n;
}

int b200orb_host_chunk_schedule(int max_pairs, int lanes, int n_pairs, int32_t* sizes, int capacity) {
    if (max_pairs < 1 || n_pairs < 1 || lanes < 1 || lanes > 2) return fail(B200ORB_E_ARG, "bad schedule arguments");
    const std::vector<int> v = host_chunk_schedule(max_pairs, lanes, n_pairs);
    if (sizes) for (int i = 0; i < (int)v.size() && i < capacity; ++i) sizes[i] = v[i];
    return (int)v.size();
}

static int batch_host_setup(b200orb_batch* b) {
    if (b->host_ready) return 0;
    const Plan& P = b->eng.hp.P;
    const size_t C = P.kp_total, PP = b->P, HW = (size_t)b->H * b->W;
    CU_TRY(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&b->s_comp, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
    const char* lv = getenv("B200ORB_HOST_LANES");
    b->lanes = lv ? std::max(1, std::min(2, atoi(lv))) : 2;
    if (b->lanes == 2) {
        CU_TRY(cudaStreamCreateWithFlags(&b->s_comp2, cudaStreamNonBlocking));
        b->eng2 = new Engine;
        b->eng2->prm = b->eng.prm; b->eng2->device = b->eng.device;
        if (b->eng2->plan(b->H, b->W, 2 * b->P) != 0) {      // no room for a second workspace: one lane, full chunks
            b->eng2->release();
            delete b->eng2;
            b->eng2 = nullptr; b->lanes = 1;
            cudaGetLastError();
            cudaStreamDestroy(b->s_comp2); b->s_comp2 = nullptr;
        }
    }
    for (int i = 0; i < b200orb_batch::NBUF; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&b->ev_in[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&b->ev_comp[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&b->ev_out[i], cudaEventDisableTiming));
        CU_TRY(cudaMalloc((void**)&b->d_in[i], 2 * PP * HW));
        CU_TRY(cudaMalloc((void**)&b->d_kps[i], 2 * PP * C * 24));
        CU_TRY(cudaMalloc((void**)&b->d_desc[i], 2 * PP * C * 32));
        CU_TRY(cudaMalloc((void**)&b->d_nkp[i], 2 * PP * 4));
        CU_TRY(cudaMalloc((void**)&b->d_uR[i], PP * C * 4));
        CU_TRY(cudaMalloc((void**)&b->d_dep[i], PP * C * 4));
        CU_TRY(cudaMalloc((void**)&b->d_mi[i], PP * C * 4));
        b->host_bytes += (long long)(2 * PP * HW + 2 * PP * C * 56 + 2 * PP * 4 + 3 * PP * C * 4);
    }
    b->host_ready = true;
    return 0;
}

int b200orb_batch_run_host(b200orb_batch* b, const uint8_t* h_left, const uint8_t* h_right, int n_pairs, double mbf, float fx,
                           float* h_kps, uint8_t* h_desc, int32_t* h_nkp, float* h_uRight, float* h_depth, int32_t* h_matchIdx) {
    return b200orb_batch_run_host_shard(b, h_left, h_right, n_pairs, mbf, fx, h_kps, h_desc, h_nkp, h_uRight, h_depth, h_matchIdx, n_pairs, 0);
}

int b200orb_batch_run_host_shard(b200orb_batch* b, const uint8_t* h_left, const uint8_t* h_right, int n_pairs, double mbf, float fx,
                                 float* h_kps, uint8_t* h_desc, int32_t* h_nkp, float* h_uRight, float* h_depth, int32_t* h_matchIdx,
                                 int job_pairs, int first_pair) {
    if (!b || !h_left || !h_right || !h_kps || !h_desc || !h_nkp || !h_uRight || !h_depth) return fail(B200ORB_E_ARG, "NULL argument");
    if (n_pairs < 1) return fail(B200ORB_E_ARG, "n_pairs must be >= 1");
    if (first_pair < 0 || job_pairs < first_pair + n_pairs) return fail(B200ORB_E_ARG, "shard [first_pair, first_pair + n_pairs) leaves the job");
    // output arrays are dimensioned for the whole job; this call fills the rows of its own pairs
    h_kps += (size_t)first_pair * b->eng.hp.P.kp_total * 6; h_desc += (size_t)first_pair * b->eng.hp.P.kp_total * 32; h_nkp += first_pair;
    h_uRight += (size_t)first_pair * b->eng.hp.P.kp_total; h_depth += (size_t)first_pair * b->eng.hp.P.kp_total;
    if (h_matchIdx) h_matchIdx += (size_t)first_pair * b->eng.hp.P.kp_total;
    CU_TRY(cudaSetDevice(b->eng.device));
    TRY(batch_host_setup(b));
    const Plan& P = b->eng.hp.P;
    const size_t C = P.kp_total, HW = (size_t)b->H * b->W, NT = (size_t)job_pairs;     // NT: pairs between the two sides of kps / desc / nkp
    if (n_pairs > b->h_status_cap) {
        if (b->h_status) cudaFreeHost(b->h_status);
        b->h_status = nullptr; b->h_status_cap = 0;
        CU_TRY(cudaHostAlloc((void**)&b->h_status, (size_t)n_pairs * sizeof(int), cudaHostAllocDefault));
        b->h_status_cap = n_pairs;
    }
    b->h_status_n = n_pairs;
    const std::vector<int> sizes = host_chunk_schedule(b->P, b->lanes, n_pairs);
    // B200ORB_HOST_TRACE=1: per-chunk completion times of upload / kernels / download on stderr (diagnostic; extra timing events)
    static const bool trace = [] { const char* v = getenv("B200ORB_HOST_TRACE"); return v && atoi(v) != 0; }();
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(3 * sizes.size() + 1);
        for (auto& e : tev) CU_TRY(cudaEventCreate(&e));
        CU_TRY(cudaEventRecord(tev[0], b->s_in));
    }
    int p0 = 0;
    for (int k = 0; k < (int)sizes.size(); p0 += sizes[k], ++k) {
        const int s = k % b200orb_batch::NBUF;
        const size_t np = (size_t)sizes[k];
        if (k >= b200orb_batch::NBUF) CU_TRY(cudaStreamWaitEvent(b->s_in, b->ev_comp[s], 0));     // inputs of chunk k-NBUF consumed
        CU_TRY(cudaMemcpyAsync(b->d_in[s], h_left + (size_t)p0 * HW, np * HW, cudaMemcpyHostToDevice, b->s_in));
        CU_TRY(cudaMemcpyAsync(b->d_in[s] + np * HW, h_right + (size_t)p0 * HW, np * HW, cudaMemcpyHostToDevice, b->s_in));
        CU_TRY(cudaEventRecord(b->ev_in[s], b->s_in));
        if (trace) CU_TRY(cudaEventRecord(tev[1 + 3 * k], b->s_in));
        const bool lane2 = b->lanes == 2 && (k & 1);
        Engine& E = lane2 ? *b->eng2 : b->eng;
        cudaStream_t sc = lane2 ? b->s_comp2 : b->s_comp;
        CU_TRY(cudaStreamWaitEvent(sc, b->ev_in[s], 0));
        if (k >= b200orb_batch::NBUF) CU_TRY(cudaStreamWaitEvent(sc, b->ev_out[s], 0));   // outputs of chunk k-NBUF downloaded
        b->status_bank = s;
        const int rrc = b->copy_only ? 0 : batch_run_on(b, E, b->d_in[s], b->d_in[s] + np * HW, (int)np, mbf, fx, b->d_kps[s], b->d_desc[s], b->d_nkp[s],
                                                         b->d_uR[s], b->d_dep[s], b->d_mi[s], sc);
        b->status_bank = 0;
        if (rrc) return rrc;
        CU_TRY(cudaEventRecord(b->ev_comp[s], sc));
        if (trace) CU_TRY(cudaEventRecord(tev[2 + 3 * k], sc));
        CU_TRY(cudaStreamWaitEvent(b->s_out, b->ev_comp[s], 0));
        // the chunk's range-error flags travel with its outputs (bank s is cleared again by chunk k + NBUF, which waits for ev_out[s])
        CU_TRY(cudaMemcpyAsync(b->h_status + p0, E.d_status + (size_t)s * E.S, np * sizeof(int), cudaMemcpyDeviceToHost, b->s_out));
        for (int side = 0; side < 2; ++side) {
            CU_TRY(cudaMemcpyAsync(h_kps + (side * NT + p0) * C * 6, b->d_kps[s] + side * np * C * 6, np * C * 24, cudaMemcpyDeviceToHost, b->s_out));
            CU_TRY(cudaMemcpyAsync(h_desc + (side * NT + p0) * C * 32, b->d_desc[s] + side * np * C * 32, np * C * 32, cudaMemcpyDeviceToHost, b->s_out));
            CU_TRY(cudaMemcpyAsync(h_nkp + side * NT + p0, b->d_nkp[s] + side * np, np * 4, cudaMemcpyDeviceToHost, b->s_out));
        }
        CU_TRY(cudaMemcpyAsync(h_uRight + (size_t)p0 * C, b->d_uR[s], np * C * 4, cudaMemcpyDeviceToHost, b->s_out));
        CU_TRY(cudaMemcpyAsync(h_depth + (size_t)p0 * C, b->d_dep[s], np * C * 4, cudaMemcpyDeviceToHost, b->s_out));
        if (h_matchIdx) CU_TRY(cudaMemcpyAsync(h_matchIdx + (size_t)p0 * C, b->d_mi[s], np * C * 4, cudaMemcpyDeviceToHost, b->s_out));
        CU_TRY(cudaEventRecord(b->ev_out[s], b->s_out));
        if (trace) CU_TRY(cudaEventRecord(tev[3 + 3 * k], b->s_out));
    }
    CU_TRY(cudaStreamSynchronize(b->s_out));
    CU_TRY(cudaStreamSynchronize(b->s_comp));
    if (b->s_comp2) CU_TRY(cudaStreamSynchronize(b->s_comp2));
    if (trace) {
        for (size_t k = 0; k < sizes.size(); ++k) {
            float t[3] = {0.f, 0.f, 0.f};
            for (int j = 0; j < 3; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[1 + 3 * k + j]);
            fprintf(stderr, "[b200orb trace] chunk %2d pairs %4d  uploaded %7.3f  computed %7.3f  downloaded %7.3f ms\n", (int)k, sizes[k], t[0], t[1], t[2]);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    for (int i = 0; i < n_pairs; ++i)
        if (b->h_status[i]) return fail(B200ORB_E_RANGE, kRangeMsg);     // every output is in host memory; flagged pairs hold -1 at the offending keypoints
    return 0;
}

// ---------------------------------------------------------------- vocabulary (BoW transform)
int b200orb_vocab_create(int n_nodes, const int32_t* child_begin, const int32_t* child_ids, const uint8_t* node_desc, int device,
                         b200orb_vocab** out) {
    if (!out) return fail(B200ORB_E_ARG, "out is NULL");
    *out = nullptr;
    if (n_nodes < 1 || !child_begin || !node_desc) return fail(B200ORB_E_ARG, "bad vocabulary arrays");
    const int nch = child_begin[n_nodes];
    if (child_begin[0] != 0 || nch < 0 || (nch > 0 && !child_ids)) return fail(B200ORB_E_ARG, "bad child CSR");
    for (int i = 0; i < n_nodes; ++i)
        if (child_begin[i + 1] < child_begin[i] || child_begin[i + 1] - child_begin[i] > 65535) return fail(B200ORB_E_ARG, "bad child CSR");
    for (int c = 0; c < nch; ++c)
        if (child_ids[c] <= 0 || child_ids[c] >= n_nodes) return fail(B200ORB_E_ARG, "child id out of range");
    CU_TRY(cudaSetDevice(device));
    b200orb_vocab* v = new b200orb_vocab;
    v->device = device; v->n_nodes = n_nodes;
    auto bail = [&](cudaError_t e, const char* what) { b200orb_vocab_destroy(v); return fail(B200ORB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&v->d_child_begin, (size_t)(n_nodes + 1) * 4)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&v->d_child_ids, (size_t)std::max(nch, 1) * 4)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&v->d_node_desc, (size_t)n_nodes * 32)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemcpy(v->d_child_begin, child_begin, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy");
    if (nch && (e = cudaMemcpy(v->d_child_ids, child_ids, (size_t)nch * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy");
    if ((e = cudaMemcpy(v->d_node_desc, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "cudaMemcpy");
    *out = v;
    return 0;
}

void b200orb_vocab_destroy(b200orb_vocab* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_child_begin); cudaFree(v->d_child_ids); cudaFree(v->d_node_desc);
    cudaFree(v->d_desc); cudaFree(v->d_leaf); cudaFree(v->d_level);
    delete v;
}

static int vocab_run(b200orb_vocab* v, const u8* d_desc, int n, int nid_level, int32_t* leaf_node, int32_t* level_node, cudaStream_t st) {
    if (n > v->cap) {
        cudaFree(v->d_leaf); cudaFree(v->d_level); cudaFree(v->d_desc);
        v->d_leaf = v->d_level = nullptr; v->d_desc = nullptr; v->cap = 0;
        CU_TRY(cudaMalloc((void**)&v->d_leaf, (size_t)n * 4));
        CU_TRY(cudaMalloc((void**)&v->d_level, (size_t)n * 4));
        CU_TRY(cudaMalloc((void**)&v->d_desc, (size_t)n * 32));
        v->cap = n;
    }
    k_vocab_descend<<<(n + VOC_WARPS - 1) / VOC_WARPS, VOC_WARPS * 32, 0, st>>>(d_desc ? d_desc : v->d_desc, n, v->d_child_begin, v->d_child_ids,
                                                                                v->d_node_desc, nid_level, v->d_leaf, v->d_level);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(leaf_node, v->d_leaf, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(level_node, v->d_level, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return 0;
}

int b200orb_vocab_transform(b200orb_vocab* v, const uint8_t* desc, int n, int nid_level, int32_t* leaf_node, int32_t* level_node) {
    if (!v || !leaf_node || !level_node || n < 0 || (n > 0 && !desc)) return fail(B200ORB_E_ARG, "bad argument");
    if (n == 0) return 0;
    CU_TRY(cudaSetDevice(v->device));
    if (n > v->cap) {       // make room first so the upload has a target
        cudaFree(v->d_leaf); cudaFree(v->d_level); cudaFree(v->d_desc);
        v->d_leaf = v->d_level = nullptr; v->d_desc = nullptr; v->cap = 0;
        CU_TRY(cudaMalloc((void**)&v->d_leaf, (size_t)n * 4));
        CU_TRY(cudaMalloc((void**)&v->d_level, (size_t)n * 4));
        CU_TRY(cudaMalloc((void**)&v->d_desc, (size_t)n * 32));
        v->cap = n;
    }
    CU_TRY(cudaMemcpy(v->d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice));
    return vocab_run(v, nullptr, n, nid_level, leaf_node, level_node, nullptr);
}

int b200orb_vocab_transform_resident(b200orb_vocab* v, b200orb_extractor* e, int nid_level, int32_t* leaf_node, int32_t* level_node) {
    if (!v || !e || !leaf_node || !level_node) return fail(B200ORB_E_ARG, "NULL argument");
    if (e->n < 0) return fail(B200ORB_E_STATE, "no extract() call yet");
    if (e->n == 0) return 0;
    if (e->eng.device != v->device) return fail(B200ORB_E_ARG, "vocabulary and extractor live on different devices");
    CU_TRY(cudaSetDevice(v->device));
    return vocab_run(v, e->d_desc, e->n, nid_level, leaf_node, level_node, e->st);
}

// ---------------------------------------------------------------- all-pairs Hamming (first piece)
int b200orb_hamming_matrix(int device, const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out) {
    if (nA < 0 || nB < 0) return fail(B200ORB_E_ARG, "negative size");
    if (nA == 0 || nB == 0) return 0;
    if (!A || !B || !out) return fail(B200ORB_E_ARG, "NULL argument");
    CU_TRY(cudaSetDevice(device));
    ScratchArena& ar = arena_of(device);
    std::lock_guard<std::mutex> guard(ar.mu);
    TRY(ar.reserve(ScratchArena::padded((size_t)nA * 32) + ScratchArena::padded((size_t)nB * 32) + ScratchArena::padded((size_t)nA * nB * 2)));
    u8* dA = ar.take<u8>((size_t)nA * 32);
    u8* dB = ar.take<u8>((size_t)nB * 32);
    unsigned short* dO = ar.take<unsigned short>((size_t)nA * nB);
    CU_TRY(cudaMemcpy(dA, A, (size_t)nA * 32, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dB, B, (size_t)nB * 32, cudaMemcpyHostToDevice));
    k_hamming_matrix<<<dim3((nB + 127) / 128, (nA + HM_ROWS - 1) / HM_ROWS), 128>>>(dA, nA, dB, nB, dO);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(out, dO, (size_t)nA * nB * 2, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------- projection searches
int b200orb_area_hamming(int device, int f32_mode, int N, const float* kxy, const int32_t* koct, const uint8_t* kdesc, int cols, int rows,
                         const int32_t* cell_start, const int32_t* cell_idx, int M, const double* qxyr, const int32_t* qlvl,
                         const int32_t* qcell, const uint8_t* qdesc, int32_t* cand_start, int32_t* cand_idx, int32_t* cand_dist, int cap,
                         int32_t* total) {
    if (N < 0 || M < 0 || cols < 1 || rows < 1 || cap < 0) return fail(B200ORB_E_ARG, "bad sizes");
    if (!cand_start || !total) return fail(B200ORB_E_ARG, "NULL argument");
    *total = 0;
    for (int q = 0; q <= M; ++q) cand_start[q] = 0;
    if (M == 0) return 0;
    if (!cell_start || !qxyr || !qlvl || !qcell || !qdesc || (N > 0 && (!kxy || !koct || !kdesc))) return fail(B200ORB_E_ARG, "NULL argument");
    const int ncell = cols * rows, nidx = cell_start[ncell];
    if (cell_start[0] != 0 || nidx < 0 || nidx > N || (nidx > 0 && !cell_idx)) return fail(B200ORB_E_ARG, "bad grid CSR");
    for (int c = 0; c < ncell; ++c) if (cell_start[c + 1] < cell_start[c]) return fail(B200ORB_E_ARG, "bad grid CSR");
    for (int i = 0; i < nidx; ++i) if (cell_idx[i] < 0 || cell_idx[i] >= N) return fail(B200ORB_E_RANGE, "grid entry is not a feature index");
    std::vector<int> qc((size_t)M * 4);
    for (int q = 0; q < M; ++q) {
        const int c0 = qcell[4 * q], c1 = qcell[4 * q + 1], r0 = qcell[4 * q + 2], r1 = qcell[4 * q + 3];
        const bool empty = c0 > c1 || r0 > r1;
        if (!empty && (c0 < 0 || c1 >= cols || r0 < 0 || r1 >= rows)) return fail(B200ORB_E_RANGE, "query cell range leaves the grid");
        qc[4 * q] = empty ? 1 : c0; qc[4 * q + 1] = empty ? 0 : c1; qc[4 * q + 2] = empty ? 0 : r0; qc[4 * q + 3] = empty ? 0 : r1;
    }
    CU_TRY(cudaSetDevice(device));
    ScratchArena& ar = arena_of(device);
    std::lock_guard<std::mutex> guard(ar.mu);
    const size_t capn = (size_t)std::max(cap, 1);
    size_t need = 0;
    for (size_t b : {(size_t)M * 24, (size_t)M * 8, (size_t)M * 16, (size_t)M * 32, (size_t)(ncell + 1) * 4, (size_t)std::max(nidx, 1) * 4,
                     (size_t)std::max(N, 1) * 8, (size_t)std::max(N, 1) * 4, (size_t)std::max(N, 1) * 32, (size_t)(M + 1) * 4, capn * 4, capn * 4})
        need += ScratchArena::padded(b);
    TRY(ar.reserve(need));
    double* d_q = ar.take<double>((size_t)M * 3);
    int* d_lvl = ar.take<int>((size_t)M * 2);
    int* d_cell = ar.take<int>((size_t)M * 4);
    u8* d_qd = ar.take<u8>((size_t)M * 32);
    int* d_cs = ar.take<int>((size_t)ncell + 1);
    int* d_ci = ar.take<int>((size_t)std::max(nidx, 1));
    float* d_xy = ar.take<float>((size_t)std::max(N, 1) * 2);
    int* d_oct = ar.take<int>((size_t)std::max(N, 1));
    u8* d_kd = ar.take<u8>((size_t)std::max(N, 1) * 32);
    int* d_cnt = ar.take<int>((size_t)M + 1);
    int* d_oi = ar.take<int>(capn);
    int* d_od = ar.take<int>(capn);
    CU_TRY(cudaMemcpy(d_q, qxyr, (size_t)M * 24, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d_lvl, qlvl, (size_t)M * 8, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d_cell, qc.data(), (size_t)M * 16, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d_qd, qdesc, (size_t)M * 32, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d_cs, cell_start, (size_t)(ncell + 1) * 4, cudaMemcpyHostToDevice));
    if (nidx) CU_TRY(cudaMemcpy(d_ci, cell_idx, (size_t)nidx * 4, cudaMemcpyHostToDevice));
    if (N) {
        CU_TRY(cudaMemcpy(d_xy, kxy, (size_t)N * 8, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_oct, koct, (size_t)N * 4, cudaMemcpyHostToDevice));
        CU_TRY(cudaMemcpy(d_kd, kdesc, (size_t)N * 32, cudaMemcpyHostToDevice));
    }
    const dim3 grid((M + AQ_WARPS - 1) / AQ_WARPS);
    auto launch = [&](int count_only, const int* d_start, int* oi, int* od) {
        if (f32_mode) k_area_hamming<float><<<grid, AQ_WARPS * 32>>>(M, d_q, d_lvl, d_cell, d_qd, rows, d_cs, d_ci, d_xy, d_oct, d_kd, count_only, d_cnt, d_start, oi, od);
        else k_area_hamming<double><<<grid, AQ_WARPS * 32>>>(M, d_q, d_lvl, d_cell, d_qd, rows, d_cs, d_ci, d_xy, d_oct, d_kd, count_only, d_cnt, d_start, oi, od);
        ++g_launches;
    };
    launch(1, nullptr, nullptr, nullptr);
    std::vector<int> cnt(M);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(cnt.data(), d_cnt, (size_t)M * 4, cudaMemcpyDeviceToHost));
    long long sum = 0;
    for (int q = 0; q < M; ++q) { cand_start[q] = (int32_t)sum; sum += cnt[q]; }
    if (sum > 0x7fffffffLL) return fail(B200ORB_E_RANGE, "too many candidates");
    cand_start[M] = (int32_t)sum; *total = (int32_t)sum;
    if (sum > cap) return fail(B200ORB_E_RANGE, "candidate buffers too small (*total holds the size needed)");
    if (sum == 0) return 0;
    if (!cand_idx || !cand_dist) return fail(B200ORB_E_ARG, "NULL argument");
    CU_TRY(cudaMemcpy(d_cnt, cand_start, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice));
    launch(0, d_cnt, d_oi, d_od);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(cand_idx, d_oi, (size_t)sum * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(cand_dist, d_od, (size_t)sum * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// The order-dependent part of the two projection searches, on the candidate lists b200orb_area_hamming produced (host code: every
// decision depends on the assignments made for the earlier map points, exactly as in the reference's loops).
//   ok[c]       0 = candidate c fails the caller-evaluated right-image check (ORBMatcher.py:246-249 / 345-349)
//   occupied[j] feature j already holds a map point with observations() > 0 (ORBMatcher.py:242-244 / 341-343); updated as matches are made
//   marks[q]    map point q has observations() > 0 (what occupied[] becomes for the feature it is assigned to)
int b200orb_greedy_project_ff(int M, const int32_t* start, const int32_t* idx, const int32_t* dist, const uint8_t* ok, int N,
                              uint8_t* occupied, const uint8_t* marks, int th_high, int32_t* best) {
    if (M < 0 || N < 0) return fail(B200ORB_E_ARG, "bad sizes");
    if (M == 0) return 0;
    if (!start || !occupied || !marks || !best || (start[M] > 0 && (!idx || !dist || !ok))) return fail(B200ORB_E_ARG, "NULL argument");
    for (int q = 0; q < M; ++q) {            // ORBMatcher.py:335-366
        int bestDist = 256, bestIdx = -1;
        for (int c = start[q]; c < start[q + 1]; ++c) {
            const int j = idx[c];
            if (j < 0 || j >= N) return fail(B200ORB_E_RANGE, "candidate index out of range");
            if (occupied[j] || !ok[c]) continue;
            if (dist[c] < bestDist) { bestDist = dist[c]; bestIdx = j; }
        }
        best[q] = -1;
        if (bestDist <= th_high && bestIdx >= 0) { best[q] = bestIdx; occupied[bestIdx] = marks[q]; }
    }
    return 0;
}
int b200orb_greedy_project_fp(int M, const int32_t* start, const int32_t* idx, const int32_t* dist, const uint8_t* ok, int N,
                              uint8_t* occupied, const uint8_t* marks, const int32_t* koct, int th_high, double nnratio, int32_t* best) {
    if (M < 0 || N < 0) return fail(B200ORB_E_ARG, "bad sizes");
    if (M == 0) return 0;
    if (!start || !occupied || !marks || !best || !koct || (start[M] > 0 && (!idx || !dist || !ok))) return fail(B200ORB_E_ARG, "NULL argument");
    for (int q = 0; q < M; ++q) {            // ORBMatcher.py:236-281
        int b1 = 256, l1 = -1, b2 = 256, l2 = -1, bi = -1;
        for (int c = start[q]; c < start[q + 1]; ++c) {
            const int j = idx[c];
            if (j < 0 || j >= N) return fail(B200ORB_E_RANGE, "candidate index out of range");
            if (occupied[j] || !ok[c]) continue;
            const int d = dist[c];
            if (d < b1) { b2 = b1; b1 = d; l2 = l1; l1 = koct[j]; bi = j; }
            else if (d < b2) { l2 = koct[j]; b2 = d; }
        }
        best[q] = -1;
        if (b1 <= th_high && bi >= 0) {
            if (l1 == l2 && (double)b1 > nnratio * (double)b2) continue;
            best[q] = bi; occupied[bi] = marks[q];
        }
    }
    return 0;
}

int b200orb_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(B200ORB_E_ARG, "p is NULL");
    CU_TRY(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return 0;
}
int b200orb_host_free(void* p) {
    CU_TRY(cudaFreeHost(p));
    return 0;
}

}  // extern "C"
