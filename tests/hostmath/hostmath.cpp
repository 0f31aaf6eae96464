// TEST INFRASTRUCTURE ONLY -- host instantiation of csrc/pixel_math.cuh.
//
// The per-pixel arithmetic of the CUDA kernels lives in __host__ __device__ functions; this file
// compiles the very same header with g++ so that tests can sweep all 2^24 colours against cv2 in a
// container without a GPU.  It is built into tests/hostmath/_build/ by tests/test_hostmath.py and
// is never loaded by the cuauv_vision_pipeline_b200 package (which fails loudly without a GPU).
#include <cstddef>
#include <cstdint>

#include "../../cuauv_vision_pipeline_b200/csrc/pixel_math.cuh"
#include "../../cuauv_vision_pipeline_b200/csrc/lab_tables.inc"

extern "C" {

// code: same values as bv_cvt_code.  x-dependent rules use `width`.
int hm_convert(const uint8_t *src, uint8_t *dst, size_t npx, int width, int code) {
    static int sdiv[256], hdiv[256];
    for (int i = 0; i < 256; ++i) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
    static int16_t luv_tab[bv::kLuvNodes * 4];
    static bool luv_built = false;
    if (code == 9 && !luv_built) {
        bv::luv_build_table(luv_tab);
        luv_built = true;
    }
    const int vec_end = width - (width % 32);
    for (size_t p = 0; p < npx; ++p) {
        const int c0 = src[3 * p], c1 = src[3 * p + 1], c2 = src[3 * p + 2];
        const bool vec = (int)(p % (size_t)width) < vec_end;
        int o0 = 0, o1 = 0, o2 = 0;
        switch (code) {
            case 0: bv::bgr2hsv(c0, c1, c2, sdiv, hdiv, o0, o1, o2); break;
            case 1: bv::bgr2lab(c0, c1, c2, kLabGammaTab, kLabCbrtTab, o0, o1, o2); break;
            case 2: dst[p] = (uint8_t)bv::bgr2gray(c0, c1, c2); continue;
            case 3: bv::bgr2ycrcb(c0, c1, c2, o0, o1, o2); break;
            case 4: bv::hsv2bgr(c0, c1, c2, vec, o0, o1, o2); break;
            case 5: bv::bgr2hls(c0, c1, c2, vec, o0, o1, o2); break;
            case 8: bv::lab2bgr(c0, c1, c2, kLabToYF, kLabInvGammaTab, o0, o1, o2); break;
            case 9: bv::bgr2luv(c0, c1, c2, luv_tab, o0, o1, o2); break;
            default: return -1;
        }
        dst[3 * p] = (uint8_t)o0;
        dst[3 * p + 1] = (uint8_t)o1;
        dst[3 * p + 2] = (uint8_t)o2;
    }
    return 0;
}

void hm_luv_table(int16_t *out) { bv::luv_build_table(out); }

void hm_hsv_tables(int *sdiv, int *hdiv) {
    for (int i = 0; i < 256; ++i) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
}

// cv2.resize INTER_LINEAR on interleaved u8 through linear_coef / linear_vblend
int hm_resize(const uint8_t *src, int sh, int sw, uint8_t *dst, int dh, int dw, int cn) {
    const double sx = (double)sw / dw, sy = (double)sh / dh;
    for (int y = 0; y < dh; ++y) {
        const bv::LinCoef cy = bv::linear_coef(y, sh, sy, false);
        for (int x = 0; x < dw; ++x) {
            const bv::LinCoef cx = bv::linear_coef(x, sw, sx, true);
            for (int c = 0; c < cn; ++c) {
                const int h0 = src[((size_t)cy.i0 * sw + cx.i0) * cn + c] * cx.w0 + src[((size_t)cy.i0 * sw + cx.i1) * cn + c] * cx.w1;
                const int h1 = src[((size_t)cy.i1 * sw + cx.i0) * cn + c] * cx.w0 + src[((size_t)cy.i1 * sw + cx.i1) * cn + c] * cx.w1;
                dst[((size_t)y * dw + x) * cn + c] = (uint8_t)bv::linear_vblend(h0, h1, cy.w0, cy.w1);
            }
        }
    }
    return 0;
}
}
