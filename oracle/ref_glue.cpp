// TEST INFRASTRUCTURE -- extern "C" front for the reference's own ORB_SLAM2::ORBextractor
// (/root/reference/pyORBExtractor/ORBextractor.{h,cpp}, compiled unmodified against oracle/cvshim).
// Mirrors what orb_extractor.cpp:22-38 + opencv_type_casters.h do on the Python boundary.
#include "ORBextractor.h"
#include <cstdint>
#include <cstdlib>
#include <new>

#ifdef REF_BUMP_ALLOC
// Monotone allocator for std::list<ExtractorNode> nodes only: under it "node address order" ==
// "node creation order", which makes the (size, pointer) sort at ORBextractor.cpp:683 reproducible
// (SURVEY.md F5).  Everything else goes to malloc.  Bound to this .so by -Wl,-Bsymbolic.
namespace {
const size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
const size_t kArenaBytes = size_t(256) << 20;
char* g_arena = nullptr;
size_t g_used = 0;
long g_overflow = 0;
}
void* operator new(size_t n) {
    if (n == kNodeBytes) {
        if (!g_arena) g_arena = (char*)std::malloc(kArenaBytes);
        size_t a = (n + 15) & ~size_t(15);
        if (g_used + a <= kArenaBytes) { void* p = g_arena + g_used; g_used += a; return p; }
        ++g_overflow;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept {
    if (g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArenaBytes) return;
    std::free(p);
}
void operator delete(void* p, size_t) noexcept { operator delete(p); }
static void arena_reset() { g_used = 0; }
extern "C" long ref_arena_overflows() { return g_overflow; }
#else
static void arena_reset() {}
extern "C" long ref_arena_overflows() { return 0; }
#endif

using ORB_SLAM2::ORBextractor;

extern "C" {
void* ref_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_destroy(void* h) { delete (ORBextractor*)h; }
// kps [cap][6] = (x, y, size, angle, response, octave) like the KeyPoint caster; desc [cap][32]
int ref_extract(void* h, const unsigned char* img, int H, int W, int cap, float* kps, unsigned char* desc) {
    arena_reset();
    ORBextractor* e = (ORBextractor*)h;
    cv::Mat image(H, W, CV_8UC1, (void*)img), mask, d;
    std::vector<cv::KeyPoint> k;
    e->operator_kd(image, mask, k, d);
    int n = (int)k.size(), m = n < cap ? n : cap;
    for (int i = 0; i < m; ++i) {
        kps[6 * i] = k[i].pt.x; kps[6 * i + 1] = k[i].pt.y; kps[6 * i + 2] = k[i].size;
        kps[6 * i + 3] = k[i].angle; kps[6 * i + 4] = k[i].response; kps[6 * i + 5] = (float)k[i].octave;
        std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    return n;
}
int ref_levels(void* h) { return ((ORBextractor*)h)->GetLevels(); }
void ref_tables(void* h, float* sf, float* isf, float* sig2, float* isig2) {
    ORBextractor* e = (ORBextractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (size_t i = 0; i < a.size(); ++i) { sf[i] = a[i]; isf[i] = b[i]; sig2[i] = c[i]; isig2[i] = d[i]; }
}
void ref_level_size(void* h, int l, int* w, int* hh) { ORBextractor* e = (ORBextractor*)h; *w = e->mvImagePyramid[l].cols; *hh = e->mvImagePyramid[l].rows; }
// what the Mat->ndarray caster hands to Python: rows*cols contiguous bytes from mat.data (opencv_type_casters.h:230-239)
void ref_level_caster_view(void* h, int l, unsigned char* out) {
    ORBextractor* e = (ORBextractor*)h;
    const cv::Mat& m = e->mvImagePyramid[l];
    std::memcpy(out, m.data, (size_t)m.rows * m.cols);
}
// the level ROI with its true row pitch (dense copy), for checking the pyramid itself
void ref_level_roi(void* h, int l, unsigned char* out) {
    ORBextractor* e = (ORBextractor*)h;
    const cv::Mat& m = e->mvImagePyramid[l];
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
}
}
