// pixel_math.cuh -- per-pixel arithmetic of the hot path, written once as __host__ __device__
// functions.  The CUDA kernels are the product; the host instantiation exists only so that
// tests/hostmath (a test-only shared object, never loaded by the package) can sweep all 2^24
// colours against cv2 in a container without a GPU.
//
// Every function reproduces, bit for bit, what OpenCV 4.x computes for 8-bit images at the
// reference's call sites (utils/color.py:11-32, modules/bins.py:13, color_balance.cpp:654,693);
// the algorithms are restated in oracle/spec_np.py and SURVEY.md Appendix A.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BV_HD __host__ __device__ __forceinline__
#else
#define BV_HD inline
#endif

// float32 operations with explicit rounding: on the device the intrinsics forbid contraction into
// FMA irrespective of compiler flags; on the host build with -ffp-contract=off.
#if defined(__CUDA_ARCH__)
#define BV_FMUL(a, b) __fmul_rn((a), (b))
#define BV_FADD(a, b) __fadd_rn((a), (b))
#define BV_FSUB(a, b) __fsub_rn((a), (b))
#define BV_FDIV(a, b) __fdiv_rn((a), (b))
#define BV_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define BV_F2I_RZ(x) __float2int_rz(x)
#define BV_F2I_RN(x) __float2int_rn(x)
#define BV_FLOOR(x) floorf(x)
#else
#define BV_FMUL(a, b) ((float)(a) * (float)(b))
#define BV_FADD(a, b) ((float)(a) + (float)(b))
#define BV_FSUB(a, b) ((float)(a) - (float)(b))
#define BV_FDIV(a, b) ((float)(a) / (float)(b))
#define BV_FMA(a, b, c) fmaf((a), (b), (c))
#define BV_F2I_RZ(x) ((int)(x))
#define BV_F2I_RN(x) ((int)nearbyintf(x))
#define BV_FLOOR(x) floorf(x)
#endif

namespace bv {

BV_HD int imin(int a, int b) { return a < b ? a : b; }
BV_HD int imax(int a, int b) { return a > b ? a : b; }
BV_HD int sat_u8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

// ------------------------------------------------------------------------------------------
// BGR -> HSV, 8-bit, H in [0,180).  sdiv[i] = rint((255<<12)/i), hdiv[i] = rint((180<<12)/(6i)),
// both 0 at i = 0 (OpenCV RGB2HSV_b, hsv_shift = 12).  Neither quotient ever lands on a tie, so
// the rounded division is (2n + i) / (2i) in integers.
// ------------------------------------------------------------------------------------------
constexpr int kHsvShift = 12;
BV_HD int hsv_sdiv(int i) { return i ? (2 * (255 << kHsvShift) + i) / (2 * i) : 0; }
BV_HD int hsv_hdiv(int i) { return i ? (2 * (180 << kHsvShift) + 6 * i) / (12 * i) : 0; }

// sdiv/hdiv: 256-entry int tables (shared memory on the device).
BV_HD void bgr2hsv(int b, int g, int r, const int *sdiv, const int *hdiv, int &h, int &s, int &v) {
    v = imax(imax(b, g), r);
    const int vmin = imin(imin(b, g), r);
    const int diff = v - vmin;
    int hh;
    if (v == r)
        hh = g - b;
    else if (v == g)
        hh = b - r + 2 * diff;
    else
        hh = r - g + 4 * diff;
    s = (diff * sdiv[v] + (1 << (kHsvShift - 1))) >> kHsvShift;
    hh = (hh * hdiv[diff] + (1 << (kHsvShift - 1))) >> kHsvShift;  // arithmetic shift
    h = hh < 0 ? hh + 180 : hh;
}

// ------------------------------------------------------------------------------------------
// HSV -> BGR, 8-bit, float32.  As cv2 4.13.0 computes it (measured over every H<180,S,V): the
// bracket is a single-rounding multiply-add; whole 32-pixel groups of a row ("vector path")
// truncate x*255, the remaining width%32 pixels of each row round to nearest even.
// ------------------------------------------------------------------------------------------
BV_HD void hsv2bgr(int H, int S, int V, bool vector_path, int &b, int &g, int &r) {
    const float hscale = 6.f / 180.f;
    const float inv255 = 1.f / 255.f;
    float h = BV_FMUL((float)H, hscale);
    const float s = BV_FMUL((float)S, inv255);
    const float v = BV_FMUL((float)V, inv255);
    float fb, fg, fr;
    if (s == 0.f) {
        fb = fg = fr = v;
    } else {
        const float fl = BV_FLOOR(h);
        int sector = (int)fl;
        const float f = BV_FSUB(h, fl);
        sector %= 6;
        if (sector < 0) sector += 6;
        const float t0 = v;
        const float t1 = BV_FMUL(v, BV_FSUB(1.f, s));
        const float t2 = BV_FMUL(v, BV_FMA(-s, f, 1.f));
        const float t3 = BV_FMUL(v, BV_FMA(-s, BV_FSUB(1.f, f), 1.f));
        // sector table {{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}} as selects
        switch (sector) {
            case 0: fb = t1; fg = t3; fr = t0; break;
            case 1: fb = t1; fg = t0; fr = t2; break;
            case 2: fb = t3; fg = t0; fr = t1; break;
            case 3: fb = t0; fg = t2; fr = t1; break;
            case 4: fb = t0; fg = t1; fr = t3; break;
            default: fb = t2; fg = t1; fr = t0; break;
        }
    }
    const float yb = BV_FMUL(fb, 255.f), yg = BV_FMUL(fg, 255.f), yr = BV_FMUL(fr, 255.f);
    if (vector_path) {
        b = sat_u8(BV_F2I_RZ(yb));
        g = sat_u8(BV_F2I_RZ(yg));
        r = sat_u8(BV_F2I_RZ(yr));
    } else {
        b = sat_u8(BV_F2I_RN(yb));
        g = sat_u8(BV_F2I_RN(yg));
        r = sat_u8(BV_F2I_RN(yr));
    }
}

// ------------------------------------------------------------------------------------------
// BGR -> Lab, 8-bit, fixed point (OpenCV RGB2Lab_b): gamma table (256 x u16), cube-root table
// (3072 x u16); the tables themselves are generated on the host in float32 (lab_tables.cpp).
// ------------------------------------------------------------------------------------------
constexpr int kLabGammaSize = 256;
constexpr int kLabCbrtSize = 3072;
BV_HD int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

BV_HD void bgr2lab(int b, int g, int r, const uint16_t *gtab, const uint16_t *ctab, int &L, int &a, int &bb) {
    const int B = gtab[b], G = gtab[g], R = gtab[r];
    const int fX = ctab[descale(R * 1777 + G * 1541 + B * 778, 12)];
    const int fY = ctab[descale(R * 871 + G * 2929 + B * 296, 12)];
    const int fZ = ctab[descale(R * 73 + G * 448 + B * 3575, 12)];
    L = sat_u8(descale(296 * fY - 1336934, 15));
    a = sat_u8(descale(500 * (fX - fY) + (128 << 15), 15));
    bb = sat_u8(descale(200 * (fY - fZ) + (128 << 15), 15));
}

// ------------------------------------------------------------------------------------------
// BGR -> GRAY (15-bit coefficients) and BGR -> YCrCb (14-bit), 8-bit.
// ------------------------------------------------------------------------------------------
BV_HD int bgr2gray(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + 16384) >> 15; }

BV_HD void bgr2ycrcb(int b, int g, int r, int &y, int &cr, int &cb) {
    y = (b * 1868 + g * 9617 + r * 4899 + 8192) >> 14;
    cr = sat_u8(((r - y) * 11682 + (128 << 14) + 8192) >> 14);
    cb = sat_u8(((b - y) * 9241 + (128 << 14) + 8192) >> 14);
}

// ------------------------------------------------------------------------------------------
// BGR -> HLS, 8-bit, float32 (OpenCV RGB2HLS_f on scaled input, then rint).  `fused` selects the
// vector formulation (multiply-add in the hue term) cv2 uses for whole 32-px groups of a row.
// ------------------------------------------------------------------------------------------
BV_HD void bgr2hls(int bi, int gi, int ri, bool fused, int &H, int &L, int &S) {
    const float inv255 = 1.f / 255.f;
    const float b = BV_FMUL((float)bi, inv255), g = BV_FMUL((float)gi, inv255), r = BV_FMUL((float)ri, inv255);
    const float vmax = fmaxf(fmaxf(b, g), r), vmin = fminf(fminf(b, g), r);
    const float diff = BV_FSUB(vmax, vmin);
    const float sm = BV_FADD(vmax, vmin);
    const float l = BV_FMUL(sm, 0.5f);
    float h = 0.f, s = 0.f;
    if (diff > 1.1920928955078125e-07f) {  // FLT_EPSILON
        s = l < 0.5f ? BV_FDIV(diff, sm) : BV_FDIV(diff, BV_FSUB(2.f, sm));
        const float k = BV_FDIV(60.f, diff);
        if (vmax == r)
            h = BV_FMUL(BV_FSUB(g, b), k);
        else if (vmax == g)
            h = fused ? BV_FMA(BV_FSUB(b, r), k, 120.f) : BV_FADD(BV_FMUL(BV_FSUB(b, r), k), 120.f);
        else
            h = fused ? BV_FMA(BV_FSUB(r, g), k, 240.f) : BV_FADD(BV_FMUL(BV_FSUB(r, g), k), 240.f);
        if (h < 0.f) h = BV_FADD(h, 360.f);
    }
    H = sat_u8(BV_F2I_RN(BV_FMUL(h, 0.5f)));
    L = sat_u8(BV_F2I_RN(BV_FMUL(l, 255.f)));
    S = sat_u8(BV_F2I_RN(BV_FMUL(s, 255.f)));
}

// ------------------------------------------------------------------------------------------
// cv2.resize INTER_LINEAR, 8-bit: 11-bit fixed-point coefficients (OpenCV resize.cpp,
// HResizeLinear / VResizeLinear for uchar).  One axis at a time.
// ------------------------------------------------------------------------------------------
struct LinCoef {
    int i0, i1;  // clamped source indices
    int w0, w1;  // int16 weights, nominally summing to 2048
};

BV_HD LinCoef linear_coef(int d, int src, double scale, bool horizontal) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)BV_FLOOR(f);
    f = BV_FSUB(f, (float)s);
    if (horizontal) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    LinCoef c;
    int w1 = BV_F2I_RN(BV_FMUL(f, 2048.f));
    int w0 = BV_F2I_RN(BV_FMUL(BV_FSUB(1.f, f), 2048.f));
    c.w1 = w1 < -32768 ? -32768 : (w1 > 32767 ? 32767 : w1);
    c.w0 = w0 < -32768 ? -32768 : (w0 > 32767 ? 32767 : w0);
    c.i0 = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
    c.i1 = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
    return c;
}

// vertical blend of two horizontally interpolated int32 rows
BV_HD int linear_vblend(int h0, int h1, int b0, int b1) {
    return sat_u8((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

}  // namespace bv
