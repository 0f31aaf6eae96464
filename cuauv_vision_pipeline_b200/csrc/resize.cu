// resize.cu -- bilinear resize (cv2.resize INTER_LINEAR on uint8: utils/transform.py:167-179,
// modules/preprocessor.py:136-143) and the YOLO input transform behind modules/yolo.py:112
// (Ultralytics LetterBox -> BGR2RGB -> HWC2CHW -> /255 -> half), batched over cameras.
//
// The interpolation is OpenCV's 11-bit fixed-point scheme (pixel_math.cuh: linear_coef /
// linear_vblend), so the uint8 result is bit-identical to cv2.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace bv {

template <int CN>
__global__ void __launch_bounds__(256) resize_kernel(const uint8_t *__restrict__ src, int sh, int sw,
                                                     uint8_t *__restrict__ dst, int dh, int dw, double scale_x,
                                                     double scale_y, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % dw);
        const size_t t = i / dw;
        const int y = (int)(t % dh);
        const size_t f = t / dh;
        const LinCoef cx = linear_coef(x, sw, scale_x, true);
        const LinCoef cy = linear_coef(y, sh, scale_y, false);
        const uint8_t *s0 = src + (f * sh + cy.i0) * (size_t)sw * CN;
        const uint8_t *s1 = src + (f * sh + cy.i1) * (size_t)sw * CN;
        uint8_t *d = dst + i * CN;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const int h0 = s0[cx.i0 * CN + c] * cx.w0 + s0[cx.i1 * CN + c] * cx.w1;
            const int h1 = s1[cx.i0 * CN + c] * cx.w0 + s1[cx.i1 * CN + c] * cx.w1;
            d[c] = (uint8_t)linear_vblend(h0, h1, cy.w0, cy.w1);
        }
    }
}

struct LetterboxImg {
    const uint8_t *src;
    int sh, sw;          // source size
    int uh, uw;          // un-padded (resized) size
    int top, left;       // padding offsets
    double scale_x, scale_y;
};

template <bool FP16>
__global__ void __launch_bounds__(256) letterbox_kernel(const LetterboxImg *__restrict__ imgs, void *__restrict__ out, int oh,
                                                        int ow, int pad) {
    const LetterboxImg im = imgs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= ow) return;
    int b = pad, g = pad, r = pad;
    const int ux = x - im.left, uy = y - im.top;
    if (ux >= 0 && ux < im.uw && uy >= 0 && uy < im.uh) {
        if (im.uw == im.sw && im.uh == im.sh) {
            const uint8_t *p = im.src + ((size_t)uy * im.sw + ux) * 3;
            b = p[0];
            g = p[1];
            r = p[2];
        } else {
            const LinCoef cx = linear_coef(ux, im.sw, im.scale_x, true);
            const LinCoef cy = linear_coef(uy, im.sh, im.scale_y, false);
            const uint8_t *s0 = im.src + (size_t)cy.i0 * im.sw * 3;
            const uint8_t *s1 = im.src + (size_t)cy.i1 * im.sw * 3;
            int v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = s0[cx.i0 * 3 + c] * cx.w0 + s0[cx.i1 * 3 + c] * cx.w1;
                const int h1 = s1[cx.i0 * 3 + c] * cx.w0 + s1[cx.i1 * 3 + c] * cx.w1;
                v[c] = linear_vblend(h0, h1, cy.w0, cy.w1);
            }
            b = v[0];
            g = v[1];
            r = v[2];
        }
    }
    // BGR -> RGB, HWC -> CHW, /255 (float32 division, then one rounding to half: what torch does
    // for `im.half() / 255`)
    const size_t plane = (size_t)oh * ow;
    const size_t o = (size_t)blockIdx.z * 3 * plane + (size_t)y * ow + x;
    const float fr = __fdiv_rn((float)r, 255.f), fg = __fdiv_rn((float)g, 255.f), fb = __fdiv_rn((float)b, 255.f);
    if (FP16) {
        __half *q = reinterpret_cast<__half *>(out);
        q[o] = __float2half_rn(fr);
        q[o + plane] = __float2half_rn(fg);
        q[o + 2 * plane] = __float2half_rn(fb);
    } else {
        float *q = reinterpret_cast<float *>(out);
        q[o] = fr;
        q[o + plane] = fg;
        q[o + 2 * plane] = fb;
    }
}

}  // namespace bv

using namespace bv;

extern "C" int bv_resize_linear(bv_ctx *ctx, const uint8_t *src_dev, int src_h, int src_w, uint8_t *dst_dev, int dst_h,
                                int dst_w, int channels, int batch) {
    BV_REQUIRE(ctx && src_dev && dst_dev, "null argument");
    BV_REQUIRE(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && batch > 0, "sizes must be positive");
    BV_REQUIRE(channels == 1 || channels == 3, "channels must be 1 or 3");
    BV_CUDA(cudaSetDevice(ctx->device));
    const size_t total = (size_t)batch * dst_h * dst_w;
    const double sx = (double)src_w / dst_w, sy = (double)src_h / dst_h;
    const int grid = grid_for(ctx, total, 256, 8);
    if (channels == 3)
        BV_LAUNCH(ctx, resize_kernel<3>, grid, 256, 0, src_dev, src_h, src_w, dst_dev, dst_h, dst_w, sx, sy, total);
    else
        BV_LAUNCH(ctx, resize_kernel<1>, grid, 256, 0, src_dev, src_h, src_w, dst_dev, dst_h, dst_w, sx, sy, total);
    return BV_OK;
}

extern "C" int bv_letterbox(bv_ctx *ctx, const uint8_t *const *srcs_host, const int32_t *heights_host,
                            const int32_t *widths_host, int n, void *out_dev, int out_h, int out_w, int pad_value,
                            int out_fp16) {
    BV_REQUIRE(ctx && srcs_host && heights_host && widths_host && out_dev, "null argument");
    BV_REQUIRE(n > 0 && n <= 65535 && out_h > 0 && out_w > 0 && out_h <= 65535, "bad batch or output size");
    BV_CUDA(cudaSetDevice(ctx->device));
    static thread_local LetterboxImg descs[256];
    BV_REQUIRE(n <= 256, "at most 256 images per call");
    for (int i = 0; i < n; ++i) {
        const int h = heights_host[i], w = widths_host[i];
        BV_REQUIRE(srcs_host[i] && h > 0 && w > 0, "bad source image");
        // Ultralytics LetterBox(auto=False, scaleup=True, center=True); python round() is
        // round-half-even == nearbyint in the default rounding mode
        const double r = fmin((double)out_h / h, (double)out_w / w);
        const int uw = (int)nearbyint(w * r), uh = (int)nearbyint(h * r);
        const double dw = (out_w - uw) / 2.0, dh = (out_h - uh) / 2.0;
        LetterboxImg &d = descs[i];
        d.src = srcs_host[i];
        d.sh = h;
        d.sw = w;
        d.uh = uh;
        d.uw = uw;
        d.top = (int)nearbyint(dh - 0.1);
        d.left = (int)nearbyint(dw - 0.1);
        d.scale_x = (double)w / uw;
        d.scale_y = (double)h / uh;
        BV_REQUIRE(uw > 0 && uh > 0, "degenerate letterbox size");
    }
    BV_TRY(ensure_scratch(ctx, SCR_LETTERBOX, sizeof(LetterboxImg) * 256));
    LetterboxImg *d_descs = (LetterboxImg *)ctx->scratch[SCR_LETTERBOX];
    BV_CUDA(cudaMemcpyAsync(d_descs, descs, sizeof(LetterboxImg) * n, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((out_w + 255) / 256, out_h, n);
    if (out_fp16)
        BV_LAUNCH(ctx, letterbox_kernel<true>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value);
    else
        BV_LAUNCH(ctx, letterbox_kernel<false>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value);
    return BV_OK;
}
