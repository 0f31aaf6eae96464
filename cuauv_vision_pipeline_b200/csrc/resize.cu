// resize.cu -- bilinear resize (cv2.resize INTER_LINEAR on uint8: utils/transform.py:167-179,
// modules/preprocessor.py:136-143) and the YOLO input transform behind modules/yolo.py:112
// (Ultralytics LetterBox -> BGR2RGB -> HWC2CHW -> /255 -> half), batched over cameras.
//
// The interpolation is OpenCV's 11-bit fixed-point scheme (pixel_math.cuh: linear_coef /
// linear_vblend), so the uint8 result is bit-identical to cv2.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

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


// ----------------------------------------------------------------------------------------------
// Letterbox with TMA-staged source rows.  A work item is one output row of one image: the two
// source rows that row interpolates between (2 x W x 3 bytes, e.g. 13 KB at 2208 px) are pulled
// into shared memory by a single elected thread with cp.async.bulk (the TMA engine; SASS: UBLKCP)
// signalling an mbarrier, so HBM sees two long sequential bursts per block instead of 12 scattered
// byte loads per output pixel; all threads then gather their taps from shared memory.
// Rows whose byte size or address is not 16-byte aligned are staged with ordinary loads instead.
// ----------------------------------------------------------------------------------------------
constexpr int kLbMaxRowBytes = 11520;  // up to 3840 px wide sources (2 stages x 2 rows = 45 KB of shared memory)

// horizontal taps of one output column
struct LbCoef {
    short o0, o1;  // byte offsets of the two taps inside a source row; o0 < 0: padding column
    short w0, w1;
};

__device__ __forceinline__ LbCoef lb_coef(const LetterboxImg &im, bool identity, int x) {
    LbCoef c;
    const int ux = x - im.left;
    if (ux < 0 || ux >= im.uw) {
        c.o0 = c.o1 = -1;
        c.w0 = c.w1 = 0;
    } else if (identity) {
        c.o0 = c.o1 = (short)(ux * 3);
        c.w0 = 2048;
        c.w1 = 0;
    } else {
        const LinCoef cx = linear_coef(ux, im.sw, im.scale_x, true);
        c.o0 = (short)(cx.i0 * 3);
        c.o1 = (short)(cx.i1 * 3);
        c.w0 = (short)cx.w0;
        c.w1 = (short)cx.w1;
    }
    return c;
}

struct LbItem {
    LetterboxImg im;
    int img, y, uy;
    bool inside_y, identity, tma_ok;
    uint32_t row_bytes;
    LinCoef cy;
    const uint8_t *g0, *g1;
};

__device__ __forceinline__ LbItem lb_item(const LetterboxImg *__restrict__ imgs, int item, int oh) {
    LbItem t;
    t.img = item / oh;
    t.y = item - t.img * oh;
    t.im = imgs[t.img];
    t.uy = t.y - t.im.top;
    t.inside_y = t.uy >= 0 && t.uy < t.im.uh;
    t.identity = t.im.uw == t.im.sw && t.im.uh == t.im.sh;
    t.row_bytes = (uint32_t)t.im.sw * 3u;
    t.cy.i0 = t.cy.i1 = t.uy;
    t.cy.w0 = 2048;
    t.cy.w1 = 0;
    if (t.inside_y && !t.identity) t.cy = linear_coef(t.uy, t.im.sh, t.im.scale_y, false);
    t.g0 = t.g1 = t.im.src;
    t.tma_ok = false;
    if (t.inside_y) {
        t.g0 = t.im.src + (size_t)t.cy.i0 * t.row_bytes;
        t.g1 = t.im.src + (size_t)t.cy.i1 * t.row_bytes;
        t.tma_ok = (t.row_bytes % 16u == 0) && ((reinterpret_cast<uintptr_t>(t.g0) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(t.g1) & 15) == 0);
    }
    return t;
}

// Persistent blocks, two-stage pipeline: while the threads interpolate output row k out of stage
// k&1, the TMA engine is already filling the other stage with the source rows of item k+1.
template <bool FP16>
__global__ void __launch_bounds__(256) letterbox_tma_kernel(const LetterboxImg *__restrict__ imgs, void *__restrict__ out, int oh,
                                                            int ow, int pad, int n_items) {
    __shared__ __align__(128) uint8_t rows[2][2][kLbMaxRowBytes];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ float norm[256];  // v / 255 in float32 (IEEE division once per value instead of per pixel)
    __shared__ LbItem sitem[2];  // the item of each stage, worked out once by thread 0 (not by all 256 threads)
    for (int k = threadIdx.x; k < 256; k += blockDim.x) norm[k] = __fdiv_rn((float)k, 255.f);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();
    auto issue = [&](int item, int stage) {  // thread 0 only
        if (item >= n_items) return;
        const LbItem t = lb_item(imgs, item, oh);
        sitem[stage] = t;
        if (!t.tma_ok) return;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done
        mbar_expect_tx(&bar[stage], t.identity ? t.row_bytes : 2u * t.row_bytes);
        bulk_g2s(rows[stage][0], t.g0, t.row_bytes, &bar[stage]);
        if (!t.identity) bulk_g2s(rows[stage][1], t.g1, t.row_bytes, &bar[stage]);
    };
    if (threadIdx.x == 0) {
        issue(blockIdx.x, 0);
        issue(blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();  // sitem[] of the first two items is visible
    uint32_t phase[2] = {0, 0};
    const size_t plane = (size_t)oh * ow;
    int k = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
        const int stage = k & 1;
        const LbItem &t = sitem[stage];
        const uint8_t *r0 = rows[stage][0];
        const uint8_t *r1 = rows[stage][1];
        if (t.inside_y) {  // block-uniform
            if (t.tma_ok) {
                mbar_wait(&bar[stage], phase[stage]);
                phase[stage] ^= 1;
            } else {  // unaligned source rows: ordinary loads
                for (uint32_t q = threadIdx.x; q < t.row_bytes; q += blockDim.x) {
                    rows[stage][0][q] = t.g0[q];
                    if (!t.identity) rows[stage][1][q] = t.g1[q];
                }
                __syncthreads();
            }
        }
        const size_t obase = (size_t)t.img * 3 * plane + (size_t)t.y * ow;
        if (t.identity) r1 = r0;  // single staged row; weights (2048, 0) make the blend an exact copy
        const bool inside_y = t.inside_y;
        const int cw0 = t.cy.w0, cw1 = t.cy.w1;
        // three output pixels per thread and step, coefficient loads issued together (independent
        // L2 round trips overlap instead of serialising over the short per-row loop)
        for (int xb = 0; xb < ow; xb += 3 * 256) {
            LbCoef cx[3];
            int xs[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                xs[u] = xb + u * 256 + threadIdx.x;
                cx[u] = lb_coef(t.im, t.identity, xs[u]);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                if (xs[u] >= ow) continue;
                int b = pad, g = pad, r = pad;
                if (inside_y && cx[u].o0 >= 0) {
                    int v[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int h0 = r0[cx[u].o0 + c] * cx[u].w0 + r0[cx[u].o1 + c] * cx[u].w1;
                        const int h1 = r1[cx[u].o0 + c] * cx[u].w0 + r1[cx[u].o1 + c] * cx[u].w1;
                        v[c] = linear_vblend(h0, h1, cw0, cw1);
                    }
                    b = v[0];
                    g = v[1];
                    r = v[2];
                }
                const float fr = norm[r], fg = norm[g], fb = norm[b];
                const int x = xs[u];
                if (FP16) {
                    __half *q = reinterpret_cast<__half *>(out);
                    q[obase + x] = __float2half_rn(fr);
                    q[obase + x + plane] = __float2half_rn(fg);
                    q[obase + x + 2 * plane] = __float2half_rn(fb);
                } else {
                    float *q = reinterpret_cast<float *>(out);
                    q[obase + x] = fr;
                    q[obase + x + plane] = fg;
                    q[obase + x + 2 * plane] = fb;
                }
            }
        }
        __syncthreads();  // everyone is done with this stage: refill it for the item after next
        if (threadIdx.x == 0) issue(item + 2 * gridDim.x, stage);
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
    bool fits = true;  // the staged kernel holds two source rows in shared memory
    for (int i = 0; i < n; ++i) fits = fits && (size_t)widths_host[i] * 3 <= (size_t)kLbMaxRowBytes;
    static const bool use_tma = getenv("BV_LETTERBOX_GATHER") == nullptr;
    if (fits && use_tma) {
        const int n_items = out_h * n;
        int grid = ctx->sm_count * 4;  // persistent: 4 blocks of 45 KB per SM
        if (grid > n_items) grid = n_items;
        if (out_fp16)
            BV_LAUNCH(ctx, letterbox_tma_kernel<true>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value, n_items);
        else
            BV_LAUNCH(ctx, letterbox_tma_kernel<false>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value, n_items);
        return BV_OK;
    }
    dim3 grid((out_w + 255) / 256, out_h, n);
    if (out_fp16)
        BV_LAUNCH(ctx, letterbox_kernel<true>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value);
    else
        BV_LAUNCH(ctx, letterbox_kernel<false>, grid, 256, 0, d_descs, out_dev, out_h, out_w, pad_value);
    return BV_OK;
}
