// hsi.cu -- the HSI contrast-correction branch of process_frame (color_balance.cpp:702-774 with
// conv_rgb_to_hsi 167-261, conv_hsi_to_rgb 263-341, percentile_min_max_qselect 144-154), P2.
//
// It runs after the RGB / HSV stages on the balanced BGR frame:
//   forward   I = (r+g+b)/3, S = 1 - min/I, H = acos((r - g/2 - b/2) / sqrt(r^2+g^2+b^2-rg-rb-gb)),
//             2 pi - H when b > g; float32 planes, clipped to [0,2pi], [0,1], [0,255] (NaN -> lower bound)
//   bounds    the 0.2 % / 99.8 % order statistics of S and of I (the reference's quickselect; its
//             std::rand() pivots do not change the value it returns) -- radix select on the device
//   backward  clip to the bounds, stretch to [0,1] / [0,255], HSI -> RGB by sector, (int) cast, clamp
// The float / double mix of every expression follows the compiled reference (restated and pinned in
// oracle/color_balance_np.py::hsi_branch: 0 differing bytes against oracle/_ref).  Tolerance (stated):
// <= 1 LSB -- acos / cos here are CUDA's double-precision functions (<= 2 ulp), the reference's are
// glibc's; a differing last bit can only show where a value sits on an integer boundary.
// Reference behaviour that is not reproduced: for frames below ~128 k pixels the second quickselect
// starts at index (int)min_value instead of 0 (line 151) and may return a neighbouring order
// statistic depending on its random pivots; here the exact order statistic is always used.
#include <math.h>

#include "balance.cuh"

namespace bv {

int select_kth_f32(bv_ctx *ctx, const float *values_dev, size_t n, size_t k, float *value_host);  // cvt.cu

__device__ __forceinline__ float clip_f(float v, float lo, float hi) {  // clip_channel_f_helper, 47-62
    if (v < lo) return lo;
    if (v > hi) return hi;
    if (v != v) return lo;
    return v;  // +-inf cannot get here: it is < lo or > hi
}

__global__ void __launch_bounds__(256) hsi_forward_kernel(const uint8_t *__restrict__ bgr, float *__restrict__ H, float *__restrict__ S,
                                                          float *__restrict__ I, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const int b = bgr[3 * p], g = bgr[3 * p + 1], r = bgr[3 * p + 2];
        const float rf = (float)r, gf = (float)g, bf = (float)b;
        const float in = (float)((double)__fadd_rn(__fadd_rn(rf, gf), bf) / 3.);              // 185
        const int mn = min(min(r, g), b);
        float s = 0.f;
        if (in > 0.f) s = (float)(1. - (double)__fdiv_rn((float)mn, in));                     // 195-200
        // 201-203: numerator in double, radicand in float
        const double num = (double)rf - (0.5 * g) - (0.5 * b);
        float t = __fmul_rn(rf, rf);
        t = __fadd_rn(t, __fmul_rn(gf, gf));
        t = __fadd_rn(t, __fmul_rn(bf, bf));
        t = __fsub_rn(t, (float)(r * g));
        t = __fsub_rn(t, (float)(r * b));
        t = __fsub_rn(t, (float)(g * b));
        float h = (float)acos(num / sqrt((double)t));
        if (b > g) h = (float)((M_PI * 2) - (double)h);                                        // 204-206
        H[p] = clip_f(h, 0.f, (float)(2. * M_PI));                                             // 255-257
        S[p] = clip_f(s, 0.f, 1.f);
        I[p] = clip_f(in, 0.f, 255.f);
    }
}

struct HsiBounds {
    float s_min, s_max, i_min, i_max, s_mult, i_mult;
};

// (unsigned char) of uchar_clip(f, 0, 255), 156-165: (int)f, then clamp.  x86's cvttss2si yields
// INT_MIN for NaN and out-of-range values, which clamps to 0.
__device__ __forceinline__ uint8_t uchar_clip(float f) {
    int n;
    if (f != f || f >= 2147483648.f || f < -2147483648.f) n = (int)0x80000000;
    else n = (int)f;
    return (uint8_t)(n < 0 ? 0 : (n > 255 ? 255 : n));
}

__device__ __forceinline__ bool feq(float a, float b) { return fabs((double)__fsub_rn(a, b)) < 0.000001; }  // 9-11

__global__ void __launch_bounds__(256) hsi_backward_kernel(const float *__restrict__ H, const float *__restrict__ S,
                                                           const float *__restrict__ I, uint8_t *__restrict__ bgr, size_t n,
                                                           HsiBounds bd) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float two_pi_3 = (float)(2. * M_PI / 3.), four_pi_3 = (float)(4. * M_PI / 3.);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const float h = H[p];
        float s = clip_f(S[p], bd.s_min, bd.s_max);                                            // 719, 725
        float i = clip_f(I[p], bd.i_min, bd.i_max);
        s = clip_f(__fmul_rn(__fsub_rn(s, bd.s_min), bd.s_mult), 0.f, 1.f);                    // 752-758
        i = clip_f(__fmul_rn(__fsub_rn(i, bd.i_min), bd.i_mult), 0.f, 255.f);
        const float is = __fmul_rn(i, s);
        const uint8_t lo = uchar_clip(__fsub_rn(i, is));                                       // i - i * s
        const uint8_t hi = uchar_clip(__fadd_rn(i, __fmul_rn(__fmul_rn(2.f, i), s)));          // i + 2 * i * s
        const double hd = (double)h, id = (double)i, isd = (double)is;
        uint8_t r, g, b;
        if (feq(h, 0.f)) {                                                                     // 277-281
            r = hi; g = lo; b = lo;
        } else if (0. < hd && hd < 2. * M_PI / 3.) {                                           // 283-287
            const double c = cos(hd) / cos(M_PI / 3. - hd);
            r = uchar_clip((float)(id + isd * c));
            g = uchar_clip((float)(id + isd * (1 - c)));
            b = lo;
        } else if (feq(h, two_pi_3)) {                                                         // 288-292
            r = lo; g = hi; b = lo;
        } else if (2. * M_PI / 3. < hd && hd < 4. * M_PI / 3.) {                               // 293-297
            const double c = cos(hd - 2. * M_PI / 3.) / cos(M_PI - hd);
            r = lo;
            g = uchar_clip((float)(id + isd * c));
            b = uchar_clip((float)(id + isd * (1 - c)));
        } else if (feq(h, four_pi_3)) {                                                        // 298-302
            r = lo; g = lo; b = hi;
        } else {                                                                               // 303-307
            const double c = cos(hd - 4. * M_PI / 3.) / cos(5. * M_PI / 3. - hd);
            r = uchar_clip((float)(id + isd * (1 - c)));
            g = lo;
            b = uchar_clip((float)(id + isd * c));
        }
        bgr[3 * p] = b;
        bgr[3 * p + 1] = g;
        bgr[3 * p + 2] = r;
    }
}

// In-place HSI contrast correction of `batch` balanced BGR frames.
int hsi_run(bv_ctx *ctx, uint8_t *bgr, int batch, size_t npx) {
    BV_TRY(ensure_scratch(ctx, SCR_HSI, sizeof(float) * 3 * npx));
    float *H = (float *)ctx->scratch[SCR_HSI], *S = H + npx, *I = S + npx;
    const int grid = grid_for(ctx, npx, 256, 8);
    const size_t lo_k = (size_t)(int)(0.002f * (float)npx), hi_k = (size_t)(int)(0.998f * (float)npx);  // 145-146
    for (int f = 0; f < batch; ++f) {
        uint8_t *frame = bgr + (size_t)f * npx * 3;
        BV_LAUNCH(ctx, hsi_forward_kernel, grid, 256, 0, frame, H, S, I, npx);
        HsiBounds bd;
        BV_TRY(select_kth_f32(ctx, S, npx, lo_k < npx ? lo_k : npx - 1, &bd.s_min));
        BV_TRY(select_kth_f32(ctx, S, npx, hi_k < npx ? hi_k : npx - 1, &bd.s_max));
        BV_TRY(select_kth_f32(ctx, I, npx, lo_k < npx ? lo_k : npx - 1, &bd.i_min));
        BV_TRY(select_kth_f32(ctx, I, npx, hi_k < npx ? hi_k : npx - 1, &bd.i_max));
        bd.s_mult = (float)(1. / (double)(bd.s_max - bd.s_min));      // 750-751: float subtraction, double division
        bd.i_mult = (float)(255. / (double)(bd.i_max - bd.i_min));
        BV_LAUNCH(ctx, hsi_backward_kernel, grid, 256, 0, H, S, I, frame, npx, bd);
    }
    return BV_OK;
}

}  // namespace bv
