// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for the handful of OpenCV C++ entry points that the reference's
// utils/color_correction/color_balance.cpp touches, so that the UNMODIFIED reference source can
// be compiled in a container that has no OpenCV C++ headers (only the python `cv2` wheel).
// Used exclusively by oracle/Makefile to build oracle/_ref/libauv-color-balance-ref.so.
//
// What the reference uses (utils/color_correction/color_balance.cpp):
//   cv::Mat(h, w, CV_8UC3, ptr)             line 369     wrap caller memory
//   cv::split / cv::merge                   371, 652, 663, 692, 696, 776
//   cv::minMaxLoc                           421-423
//   cv::mean(...).val[0]                    426-428
//   cv::cvtColor BGR2HSV / HSV2BGR          654, 693
//
// The two 8-bit colour conversions follow OpenCV 4.x's published integer / float32 algorithms
// (SURVEY.md A.1, A.2); tests/test_oracle_ref.py proves the compiled result equal to a
// restatement that calls the real cv2.cvtColor, so this shim is pinned by cv2 4.13.0 itself.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)

namespace cv {

enum ColorConversionCodes { COLOR_BGR2HSV = 40, COLOR_HSV2BGR = 54 };

struct Point {
    int x = 0, y = 0;
};

struct Scalar {
    double val[4] = {0, 0, 0, 0};
    double operator[](int i) const { return val[i]; }
};

class Mat {
  public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr;

    Mat() = default;
    // Wraps caller-owned memory; never frees it.
    Mat(int r, int c, int type, void *ptr) : rows(r), cols(c), data((unsigned char *)ptr), cn_(1 + (type >> 3)) {}
    Mat(size_t r, size_t c, int type, void *ptr) : Mat((int)r, (int)c, type, ptr) {}

    int channels() const { return cn_; }
    size_t total() const { return (size_t)rows * (size_t)cols; }
    bool empty() const { return data == nullptr; }

    // OpenCV semantics: a no-op when the shape and type already match (so writing "into" a Mat
    // that wraps the caller's buffer really lands in that buffer), else allocate fresh storage.
    void create(int r, int c, int type) {
        int cn = 1 + (type >> 3);
        if (data && r == rows && c == cols && cn == cn_) return;
        rows = r;
        cols = c;
        cn_ = cn;
        store_ = std::shared_ptr<unsigned char>((unsigned char *)std::malloc((size_t)r * c * cn), std::free);
        data = store_.get();
    }

  private:
    int cn_ = 1;
    std::shared_ptr<unsigned char> store_;
};

inline void split(const Mat &src, Mat *planes) {
    const int cn = src.channels();
    const size_t n = src.total();
    for (int k = 0; k < cn; ++k) planes[k].create(src.rows, src.cols, CV_8UC1);
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < cn; ++k) planes[k].data[i] = src.data[i * cn + k];
}

inline void merge(const Mat *planes, size_t count, Mat &dst) {
    const int cn = (int)count;
    dst.create(planes[0].rows, planes[0].cols, CV_MAKETYPE(CV_8U, cn));
    const size_t n = planes[0].total();
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < cn; ++k) dst.data[i * cn + k] = planes[k].data[i];
}

// cv::mean on 8-bit data: exact integer sum, one division in double.
inline Scalar mean(const Mat &m) {
    Scalar s;
    const int cn = m.channels();
    const size_t n = m.total();
    uint64_t acc[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < cn; ++k) acc[k] += m.data[i * cn + k];
    for (int k = 0; k < cn; ++k) s.val[k] = n ? (double)acc[k] / (double)n : 0.0;
    return s;
}

inline void minMaxLoc(const Mat &m, double *minv, double *maxv, Point *, Point *) {
    const size_t n = m.total() * m.channels();
    unsigned char lo = 255, hi = 0;
    for (size_t i = 0; i < n; ++i) {
        if (m.data[i] < lo) lo = m.data[i];
        if (m.data[i] > hi) hi = m.data[i];
    }
    if (minv) *minv = lo;
    if (maxv) *maxv = hi;
}

namespace shim_detail {

struct HsvTables {
    int sdiv[256], hdiv[256];
    HsvTables() {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; ++i) {
            sdiv[i] = (int)std::nearbyint((255 << 12) / (1. * i));
            hdiv[i] = (int)std::nearbyint((180 << 12) / (6. * i));
        }
    }
};

// 8-bit BGR -> HSV, hue range 180 (integer tables, 12-bit shift).
inline void bgr2hsv_row(const unsigned char *src, unsigned char *dst, size_t n) {
    static const HsvTables t;
    for (size_t i = 0; i < n; ++i, src += 3, dst += 3) {
        int b = src[0], g = src[1], r = src[2];
        int v = b > g ? b : g;
        v = v > r ? v : r;
        int vmin = b < g ? b : g;
        vmin = vmin < r ? vmin : r;
        int diff = v - vmin;
        int h;
        if (v == r)
            h = g - b;
        else if (v == g)
            h = b - r + 2 * diff;
        else
            h = r - g + 4 * diff;
        int s = (diff * t.sdiv[v] + (1 << 11)) >> 12;
        h = (h * t.hdiv[diff] + (1 << 11)) >> 12;
        if (h < 0) h += 180;
        dst[0] = (unsigned char)h;
        dst[1] = (unsigned char)s;
        dst[2] = (unsigned char)v;
    }
}

// 8-bit HSV -> BGR in float32 as cv2 4.13.0 computes it (measured over every H<180,S,V, see
// tests/test_oracle_spec.py): single-rounding multiply-add inside the bracket everywhere; the
// vector path (whole 32-pixel groups of a row) TRUNCATES x*255, the row tail (width mod 32
// pixels) ROUNDS to nearest-even (saturate_cast).
inline void hsv2bgr_px(const unsigned char *src, unsigned char *dst, bool vector_path) {
    static const int sector_tab[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    float h = (float)src[0] * (6.f / 180.f);
    float s = (float)src[1] * (1.f / 255.f);
    float v = (float)src[2] * (1.f / 255.f);
    float b, g, r;
    if (s == 0.f) {
        b = g = r = v;
    } else {
        float fl = std::floor(h);
        int sector = (int)fl;
        float f = h - fl;
        sector %= 6;
        if (sector < 0) sector += 6;
        float tab[4];
        tab[0] = v;
        tab[1] = v * (1.f - s);
        tab[2] = v * std::fmaf(-s, f, 1.f);
        tab[3] = v * std::fmaf(-s, 1.f - f, 1.f);
        b = tab[sector_tab[sector][0]];
        g = tab[sector_tab[sector][1]];
        r = tab[sector_tab[sector][2]];
    }
    auto to8 = [vector_path](float x) -> unsigned char {
        float y = x * 255.f;
        int t = vector_path ? (int)y : (int)std::nearbyintf(y);
        return (unsigned char)(t < 0 ? 0 : (t > 255 ? 255 : t));
    };
    dst[0] = to8(b);
    dst[1] = to8(g);
    dst[2] = to8(r);
}

}  // namespace shim_detail

}  // namespace cv

// Optional delegate for cv::cvtColor: when a host program (bench.py's reference arm) installs a function here, the two
// conversions run in REAL OpenCV (cv2.cvtColor on views of the same buffers, SIMD and all) instead of the scalar formulas
// below.  The parity tests never install it; the bench asserts that both ways give identical bytes before timing.
extern "C" {
typedef void (*bv_shim_cvtcolor_hook_t)(const unsigned char *src, unsigned char *dst, int rows, int cols, int code);
__attribute__((visibility("default"))) bv_shim_cvtcolor_hook_t bv_shim_cvtcolor_hook = nullptr;
}

namespace cv {

inline void cvtColor(const Mat &src, Mat &dst, int code) {
    // src may alias dst only when the conversion is per-pixel in place, which holds here.
    Mat out = dst;
    out.create(src.rows, src.cols, CV_8UC3);
    if (bv_shim_cvtcolor_hook && (code == COLOR_BGR2HSV || code == COLOR_HSV2BGR)) {
        bv_shim_cvtcolor_hook(src.data, out.data, src.rows, src.cols, code);
        dst = out;
        return;
    }
    const size_t W = (size_t)src.cols;
    for (int y = 0; y < src.rows; ++y) {
        const unsigned char *s = src.data + (size_t)y * W * 3;
        unsigned char *d = out.data + (size_t)y * W * 3;
        if (code == COLOR_BGR2HSV) {
            shim_detail::bgr2hsv_row(s, d, W);
        } else if (code == COLOR_HSV2BGR) {
            const size_t vec_end = W - (W % 32);
            for (size_t x = 0; x < W; ++x) shim_detail::hsv2bgr_px(s + 3 * x, d + 3 * x, x < vec_end);
        } else {
            std::abort();
        }
    }
    dst = out;
}

}  // namespace cv
