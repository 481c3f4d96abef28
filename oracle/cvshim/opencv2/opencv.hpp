// TEST INFRASTRUCTURE -- a header-only stand-in for the slice of the OpenCV 4 C++ API that the
// reference extractor uses, so that /root/reference/pyORBExtractor/ORBextractor.cpp can be compiled
// UNMODIFIED in a container that has no OpenCV C++ SDK.  It is NOT OpenCV: the
// pixel primitives forward to oracle/cvprims.hpp, each of which is pinned bit-for-bit against
// cv2 4.13.0 by tests/test_oracle_prims.py.  Only the members the reference touches exist.
#pragma once
#include <cassert>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>
#include <list>
#include <iterator>
#include <cstdlib>
#include <algorithm>
#include "../../cvprims.hpp"

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5

inline int cvRound(double v) { return orbo::cv_round(v); }
inline int cvRound(float v) { return orbo::cv_round(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return orbo::cv_floor(v); }
inline int cvCeil(double v) { return orbo::cv_ceil(v); }

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

struct MatZeros { int rows, cols, type; };

class Mat {
public:
    int rows, cols; uchar* data; size_t step;
    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(Size sz, int type) { alloc(sz.height, sz.width); (void)type; }
    Mat(int r, int c, int type) { alloc(r, c); (void)type; }
    Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((uchar*)ext), step((size_t)c) { (void)type; }
    void create(int r, int c, int type) { (void)type; if (r != rows || c != cols || !data) alloc(r, c); }
    static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
    Mat& operator=(const MatZeros& z) {   // MatExpr assignment: reuse the buffer when the shape already matches
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const { Mat m(rows, cols, 0); for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols); return m; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    size_t step1() const { return step; }
    int type() const { return CV_8UC1; }
    int depth() const { return CV_8U; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { owner.reset(); data = nullptr; rows = cols = 0; step = 0; }
private:
    std::shared_ptr<uchar> owner;
    void alloc(int r, int c) {
        rows = r; cols = c; step = (size_t)c;
        owner.reset((uchar*)std::malloc((size_t)r * c + 64), std::free);
        data = owner.get();
    }
};

class _InputArray {
    const Mat* m;
public:
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return m->empty(); }
    Mat getMat() const { return *m; }
};
class _OutputArray {
    Mat* m;
public:
    _OutputArray(Mat& mm) : m(&mm) {}
    void release() const { m->release(); }
    void create(int r, int c, int type) const { m->create(r, c, type); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline float fastAtan2(float y, float x) { return orbo::fast_atan2_deg(y, x); }

inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nonmax = true) {
    assert(nonmax);
    std::vector<orbo::FastKp> out;
    orbo::fast9_16_nms(img.data, img.cols, img.rows, img.step, threshold, out);
    kps.clear();
    for (const orbo::FastKp& k : out) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}

inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int interp) {
    assert(interp == INTER_LINEAR);
    dst.create(sz.height, sz.width, src.type());
    orbo::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

inline void copyMakeBorder(const Mat& src, Mat& dst, int t, int b, int l, int r, int type) {
    assert(t == b && l == r && t == l && (type & 7) == BORDER_REFLECT_101);
    // every use in the reference reflects the given ROI itself (level 0: a standalone image; others: BORDER_ISOLATED)
    dst.create(src.rows + 2 * t, src.cols + 2 * t, src.type());
    orbo::copy_make_border_101(src.data, src.cols, src.rows, src.step, dst.data, dst.step, t);
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size k, double sx, double sy, int border) {
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2 && border == BORDER_REFLECT_101);
    Mat tmp = src.clone();
    dst.create(src.rows, src.cols, src.type());
    orbo::gaussian_blur7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}

struct KeyPointsFilter {   // only reached from the dead ComputeKeyPointsOld; must link
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            k.resize(n);
        }
    }
};

}  // namespace cv
