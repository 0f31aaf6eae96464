// balance.cu -- underwater colour balance on the device.  Replaces the reference's process_frame
// (utils/color_correction/color_balance.cpp:343-780) for every flag combination except the HSI
// branch (702-774) and tiled equalisation (horizontal/vertical_blocks > 1).
//
// The reference makes ~14 full passes over split planes on the CPU.  Here the whole algorithm is
// three streaming passes over the interleaved frame plus two tiny per-frame statistics kernels:
//
//   pass 1  hist_bgr    256-bin histograms of B, G, R                      (reads 3 B/px)
//   stats1  percentile bounds (112-142), exact clipped means from the histogram (426-428),
//           dominant channel + gains in double (480-544), optional RGB contrast stretch
//           (546-645) -> ONE composed 256-entry table per channel
//   pass 2  hist_sv     table -> BGR2HSV -> histograms of S and V          (re-reads 3 B/px, L2)
//   stats2  S/V percentile bounds (671-681) -> stretch tables (683-686)
//   pass 3  final       table -> BGR2HSV -> S/V tables -> HSV2BGR -> [convert -> inRange]
//                       writes balanced BGR and/or converted image and/or mask
//
// Because clipping, equalisation and the contrast stretch are all per-channel point operations
// that depend only on global statistics, they collapse into look-up tables computed once per frame;
// the exact integer channel sums come out of the histogram, so no extra reduction pass is needed.
// Frames are processed in chunks sized to stay L2-resident so that passes 2 and 3 do not touch HBM.
#include <stdlib.h>

#include "balance.cuh"
#include "convert.cuh"

namespace bv {

constexpr int kBalThreads = 256;
constexpr int kBalWarps = kBalThreads / 32;

// ----------------------------------------------------------------------------------------------
// pass 1: BGR histograms.  One private 3x256 histogram per warp in shared memory, merged into
// the frame's global histogram with one atomic per non-empty bin per block.
// ----------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kBalThreads) hist_bgr_kernel(const uint8_t *__restrict__ src, BalFrame *__restrict__ st,
                                                               size_t npx) {
    __shared__ uint32_t h[kBalWarps][3][256];
    for (int i = threadIdx.x; i < kBalWarps * 768; i += blockDim.x) (&h[0][0][0])[i] = 0;
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t(*hw)[256] = h[threadIdx.x >> 5];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t ngroups = VEC ? npx / 16 : 0;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        Px16 in;
        load_px16<true>(f, g, in);
#pragma unroll
        for (int k = 0; k < 48; ++k) atomicAdd(&hw[k % 3][BV_GETB(in.w, k)], 1u);
    }
    for (size_t p = ngroups * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
        atomicAdd(&hw[0][f[3 * p]], 1u);
        atomicAdd(&hw[1][f[3 * p + 1]], 1u);
        atomicAdd(&hw[2][f[3 * p + 2]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kBalWarps; ++w) s += (&h[w][0][0])[i];
        if (s) atomicAdd(&st[frame].hist_bgr[0][0] + i, s);
    }
}

// percentile_min_max (color_balance.cpp:112-142) on a finished histogram
__device__ void percentile_bounds(const uint32_t *cnt, size_t n, int &lo, int &hi) {
    // float32 products truncated to int, exactly as lines 113-114
    int low_bound = (int)__fmul_rn(0.002f, (float)n);
    int high_bound = (int)(n - (size_t)(int)__fmul_rn(0.998f, (float)n));
    lo = 0;
    hi = 255;
    for (int i = 0; i < 256; ++i) {
        if (low_bound < (int)cnt[i]) {
            lo = i;
            break;
        }
        low_bound -= (int)cnt[i];
    }
    for (int i = 255; i >= 0; --i) {
        if (high_bound < (int)cnt[i]) {
            hi = i;
            break;
        }
        high_bound -= (int)cnt[i];
    }
}

__device__ void extrema_bounds(const uint32_t *cnt, int &lo, int &hi) {  // cv::minMaxLoc, 421-423
    lo = 0;
    hi = 255;
    for (int i = 0; i < 256; ++i)
        if (cnt[i]) {
            lo = i;
            break;
        }
    for (int i = 255; i >= 0; --i)
        if (cnt[i]) {
            hi = i;
            break;
        }
}

// constrain(val, 0, 255) of color_balance.cpp:13-23: clamp in double, truncate.  NaN (0 * inf when
// a channel mean is 0) is undefined behaviour in the reference; defined here as 0.
__device__ __forceinline__ int constrain255(double v) {
    if (v < 0.0) return 0;
    if (v > 255.0) return 255;
    if (v != v) return 0;
    return (int)v;
}

// `(unsigned char)double` as the compiled reference does it on x86-64 (truncate to int, keep the
// low byte); out-of-range input is undefined behaviour in the reference (634-640).
__device__ __forceinline__ int uchar_cast(double v) {
    if (v != v || v >= 2147483648.0 || v <= -2147483649.0) return 0;
    return ((int)v) & 0xFF;
}

// ----------------------------------------------------------------------------------------------
// stats1: one block per frame
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_bgr_kernel(BalFrame *__restrict__ st, size_t npx, bv_balance_params prm,
                                                        const double *__restrict__ pow_quarter) {
    BalFrame &F = st[blockIdx.x];
    __shared__ int s_lo[3], s_hi[3];
    __shared__ double s_avg[3], s_gain[3];
    __shared__ double s_ratio[3];      // per channel ratio for the rgb contrast stretch
    __shared__ int s_cc_min[3];
    const int t = threadIdx.x;
    if (t < 3) {  // channel t (0=B 1=G 2=R); the three channels are independent
        int lo, hi;
        if (prm.rgb_extrema_clipping)
            percentile_bounds(F.hist_bgr[t], npx, lo, hi);
        else
            extrema_bounds(F.hist_bgr[t], lo, hi);
        unsigned long long sum = 0;  // exact sum of the clipped channel == cv::mean numerator
        for (int i = 0; i < 256; ++i) {
            const int c = i < lo ? lo : (i > hi ? hi : i);
            sum += (unsigned long long)c * F.hist_bgr[t][i];
        }
        s_lo[t] = lo;
        s_hi[t] = hi;
        s_avg[t] = (double)sum / (double)npx;
    }
    __syncthreads();
    if (t == 0) {
        const double b = s_avg[0], g = s_avg[1], r = s_avg[2];
        int dom;
        // 480 / 501 / 522: red if strictly largest, else green if strictly largest, else blue
        if (r > g && r > b) dom = 2;
        else if (g > r && g > b) dom = 1;
        else dom = 0;
        for (int c = 0; c < 3; ++c) s_gain[c] = (c == dom) ? 1.0 : s_avg[dom] / s_avg[c];
        F.stats.dominant = dom;
        for (int c = 0; c < 3; ++c) {
            F.stats.bgr_min[c] = s_lo[c];
            F.stats.bgr_max[c] = s_hi[c];
            F.stats.bgr_avg[c] = s_avg[c];
        }
        F.stats.s_min = F.stats.v_min = 0;
        F.stats.s_max = F.stats.v_max = 255;
        F.stats.degenerate = 0;
        if (prm.rgb_contrast_correct) {
            // 560-627: order channels by the (pre-equalisation) means
            int mx, md, mn;
            if (r > g) {
                if (r > b) { mx = 2; if (g > b) { md = 1; mn = 0; } else { md = 0; mn = 1; } }
                else { mx = 0; md = 2; mn = 1; }
            } else {
                if (g > b) { mx = 1; if (r > b) { md = 2; mn = 0; } else { md = 0; mn = 2; } }
                else { mx = 0; md = 1; mn = 2; }
            }
            const double desired_max = (double)((s_hi[mn] + s_hi[md] + s_hi[mx]) / 3);   // 629: int division
            s_ratio[mn] = (desired_max - s_lo[mn]) / (double)(s_hi[mn] - s_lo[mn]);
            s_ratio[md] = (desired_max - 0.0) / (double)(s_hi[md] - s_lo[md]);
            s_ratio[mx] = ((double)s_hi[mx] - 0.0) / (double)(s_hi[mx] - s_lo[mx]);
            for (int c = 0; c < 3; ++c) s_cc_min[c] = s_lo[c];
        }
    }
    __syncthreads();
    // every thread builds entry t of the three composed tables
    for (int c = 0; c < 3; ++c) {
        int x = t < s_lo[c] ? s_lo[c] : (t > s_hi[c] ? s_hi[c] : t);                 // clip_channel, 25-45
        if (prm.equalize_rgb && c != F.stats.dominant) {
            const double xd = (double)x;
            if (prm.adaptive_cast_correction)                                        // 489-491
                x = constrain255(xd * (pow_quarter[x] * (s_gain[c] - 1.) + 1.));
            else                                                                     // 494-495
                x = constrain255(xd * s_gain[c]);
        }
        if (prm.rgb_contrast_correct) x = uchar_cast((double)(x - s_cc_min[c]) * s_ratio[c]);  // 634-640
        F.lut_bgr[c][t] = (uint8_t)x;
    }
}

// ----------------------------------------------------------------------------------------------
// pass 2: S and V histograms of the table-corrected frame
// ----------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kBalThreads) hist_sv_kernel(const uint8_t *__restrict__ src, BalFrame *__restrict__ st,
                                                              size_t npx) {
    __shared__ uint32_t h[kBalWarps][2][256];
    __shared__ uint8_t lut[3][256];
    __shared__ int sdiv[256], hdiv[256];
    const int frame = blockIdx.y;
    for (int i = threadIdx.x; i < kBalWarps * 512; i += blockDim.x) (&h[0][0][0])[i] = 0;
    for (int i = threadIdx.x; i < 768; i += blockDim.x) (&lut[0][0])[i] = (&st[frame].lut_bgr[0][0])[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = hsv_sdiv(i);
        hdiv[i] = hsv_hdiv(i);
    }
    __syncthreads();
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t(*hw)[256] = h[threadIdx.x >> 5];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t ngroups = VEC ? npx / 16 : 0;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        Px16 in;
        load_px16<true>(f, g, in);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int hh, ss, vv;
            bgr2hsv(lut[0][BV_GETB(in.w, 3 * j)], lut[1][BV_GETB(in.w, 3 * j + 1)], lut[2][BV_GETB(in.w, 3 * j + 2)], sdiv,
                    hdiv, hh, ss, vv);
            atomicAdd(&hw[0][ss], 1u);
            atomicAdd(&hw[1][vv], 1u);
        }
    }
    for (size_t p = ngroups * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
        int hh, ss, vv;
        bgr2hsv(lut[0][f[3 * p]], lut[1][f[3 * p + 1]], lut[2][f[3 * p + 2]], sdiv, hdiv, hh, ss, vv);
        atomicAdd(&hw[0][ss], 1u);
        atomicAdd(&hw[1][vv], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kBalWarps; ++w) s += (&h[w][0][0])[i];
        if (s) atomicAdd(&st[frame].hist_sv[0][0] + i, s);
    }
}

// stats2: S/V percentile bounds and stretch tables (671-686)
__global__ void __launch_bounds__(256) stats_sv_kernel(BalFrame *__restrict__ st, size_t npx) {
    BalFrame &F = st[blockIdx.x];
    __shared__ int lo[2], hi[2];
    const int t = threadIdx.x;
    if (t < 2) percentile_bounds(F.hist_sv[t], npx, lo[t], hi[t]);
    __syncthreads();
    if (t == 0) {
        F.stats.s_min = lo[0];
        F.stats.s_max = hi[0];
        F.stats.v_min = lo[1];
        F.stats.v_max = hi[1];
        F.stats.degenerate = (lo[0] == hi[0] || lo[1] == hi[1]) ? 1 : 0;
    }
    for (int c = 0; c < 2; ++c) {
        const int x = t < lo[c] ? lo[c] : (t > hi[c] ? hi[c] : t);
        // int32, C division.  The reference divides by zero (SIGFPE) when hi == lo; defined as 0 here.
        const int d = hi[c] - lo[c];
        F.lut_sv[c][t] = (uint8_t)(d ? ((x - lo[c]) * 255) / d : 0);
    }
}

// ----------------------------------------------------------------------------------------------
// pass 3: everything per pixel.  MODE 0: no balance (pure conversion); 1: BGR tables only
// (hsv_contrast_correct = 0); 2: tables + HSV stretch round trip.  CODE: conversion applied to
// the balanced pixel for the converted / mask outputs (-1: none).
// ----------------------------------------------------------------------------------------------
struct FinalSmem {
    uint8_t lut[3][256];
    uint8_t lut_sv[2][256];
    int sdiv[256], hdiv[256];
};

template <int MODE>
__device__ __forceinline__ void balance_px(int &b, int &g, int &r, bool vec, const FinalSmem &fs) {
    if (MODE >= 1) {
        b = fs.lut[0][b];
        g = fs.lut[1][g];
        r = fs.lut[2][r];
    }
    if (MODE == 2) {
        int h, s, v;
        bgr2hsv(b, g, r, fs.sdiv, fs.hdiv, h, s, v);
        hsv2bgr(h, fs.lut_sv[0][s], fs.lut_sv[1][v], vec, b, g, r);
    }
}

template <int MODE, int CODE, bool VEC>
__global__ void __launch_bounds__(kBalThreads) final_kernel(const uint8_t *__restrict__ src, const BalFrame *__restrict__ st,
                                                            size_t npx, int width, BalOutputs out,
                                                            const uint16_t *__restrict__ g_gamma,
                                                            const uint16_t *__restrict__ g_cbrt) {
    __shared__ FinalSmem fs;
    __shared__ SmemTabs tabs;
    const int frame = blockIdx.y;
    if (MODE >= 1)
        for (int i = threadIdx.x; i < 768; i += blockDim.x) (&fs.lut[0][0])[i] = (&st[frame].lut_bgr[0][0])[i];
    if (MODE == 2) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) (&fs.lut_sv[0][0])[i] = (&st[frame].lut_sv[0][0])[i];
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            fs.sdiv[i] = hsv_sdiv(i);
            fs.hdiv[i] = hsv_hdiv(i);
        }
    }
    init_tabs<CODE>(tabs, g_gamma, g_cbrt);  // ends with __syncthreads()
    const size_t foff = (size_t)frame * npx;
    const uint8_t *f = src + foff * 3;
    const int vec_end = width - (width % 32);
    constexpr bool kOne = CvtTraits<CODE>::kOneChannel;
    constexpr bool kNeedX = (MODE == 2) || CvtTraits<CODE>::kNeedsX;
    Bounds3 bd;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        bd.lo[k] = out.lo[k];
        bd.hi[k] = out.hi[k];
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t ngroups = VEC ? npx / 16 : 0;
    const int wp2 = ((width + 31) / 32) * 2;  // uint16 units per bit-packed row
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        Px16 in;
        load_px16<false>(f, g, in);
        const size_t p0 = g * 16;
        const int y = (int)(p0 / (size_t)width);
        const int x0 = (int)(p0 - (size_t)y * width);
        int x = x0;
        Px16 ob, oc;
        uint32_t q[4] = {0, 0, 0, 0};
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) ob.w[k] = oc.w[k] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int b = BV_GETB(in.w, 3 * j), gg = BV_GETB(in.w, 3 * j + 1), r = BV_GETB(in.w, 3 * j + 2);
            const bool vec = x < vec_end;
            balance_px<MODE>(b, gg, r, vec, fs);
            if (kNeedX) {
                if (++x == width) x = 0;
            }
            BV_PUTB(ob.w, 3 * j, b);
            BV_PUTB(ob.w, 3 * j + 1, gg);
            BV_PUTB(ob.w, 3 * j + 2, r);
            {
                int o0, o1, o2;
                convert_px<CODE>(b, gg, r, vec, tabs, o0, o1, o2);
                if (kOne) {
                    BV_PUTB(q, j, o0);
                } else {
                    BV_PUTB(oc.w, 3 * j, o0);
                    BV_PUTB(oc.w, 3 * j + 1, o1);
                    BV_PUTB(oc.w, 3 * j + 2, o2);
                }
                if (in_range_px<CODE>(o0, o1, o2, bd)) bits |= 1u << j;
            }
        }
        if (out.balanced) store_px16(out.balanced + foff * 3, g, ob);
        if (out.converted) {
            if (kOne)
                st_stream(reinterpret_cast<uint4 *>(out.converted + foff) + g, make_uint4(q[0], q[1], q[2], q[3]));
            else
                store_px16(out.converted + foff * 3, g, oc);
        }
        if (out.mask) {
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t nib = (bits >> (4 * k)) & 0xF;
                // spread 4 bits to 4 bytes of 0x00 / 0xFF
                m[k] = ((nib & 1) * 0xFFu) | (((nib >> 1) & 1) * 0xFF00u) | (((nib >> 2) & 1) * 0xFF0000u) |
                       (((nib >> 3) & 1) * 0xFF000000u);
            }
            st_stream(reinterpret_cast<uint4 *>(out.mask + foff) + g, make_uint4(m[0], m[1], m[2], m[3]));
        }
        if (out.mask_bits)  // requires width % 16 == 0: a group never straddles rows
            out.mask_bits[(size_t)frame * wp2 * (npx / (size_t)width) + (size_t)y * wp2 + (x0 >> 4)] = (uint16_t)bits;
    }
    // scalar path: trailing pixels of each frame, or everything for unaligned / odd-sized frames
    for (size_t p = ngroups * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
        const int x = (int)(p % (size_t)width);
        int b = f[3 * p], gg = f[3 * p + 1], r = f[3 * p + 2];
        const bool vec = x < vec_end;
        balance_px<MODE>(b, gg, r, vec, fs);
        if (out.balanced) {
            uint8_t *o = out.balanced + (foff + p) * 3;
            o[0] = (uint8_t)b;
            o[1] = (uint8_t)gg;
            o[2] = (uint8_t)r;
        }
        {
            int o0, o1, o2;
            convert_px<CODE>(b, gg, r, vec, tabs, o0, o1, o2);
            if (out.converted) {
                if (kOne) {
                    out.converted[foff + p] = (uint8_t)o0;
                } else {
                    uint8_t *o = out.converted + (foff + p) * 3;
                    o[0] = (uint8_t)o0;
                    o[1] = (uint8_t)o1;
                    o[2] = (uint8_t)o2;
                }
            }
            if (out.mask) out.mask[foff + p] = in_range_px<CODE>(o0, o1, o2, bd) ? 255 : 0;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static bool vec_ok(const void *p, size_t npx, int batch, int bytes_per_px) {
    (void)bytes_per_px;
    return p == nullptr || (host_aligned16(p) && (npx % 16 == 0 || batch == 1));
}

template <int MODE, int CODE>
static int launch_final(bv_ctx *ctx, const uint8_t *src, const BalFrame *st, int batch, size_t npx, int width,
                        const BalOutputs &out, bool vec) {
    int bpf = (ctx->sm_count * 4 + batch - 1) / batch;
    const size_t need = (npx / 16 + kBalThreads - 1) / kBalThreads;
    if ((size_t)bpf > need) bpf = (int)(need ? need : 1);
    dim3 grid(bpf, batch);
    if (vec)
        BV_LAUNCH(ctx, (final_kernel<MODE, CODE, true>), grid, kBalThreads, 0, src, st, npx, width, out, ctx->d_lab_gamma,
                  ctx->d_lab_cbrt);
    else
        BV_LAUNCH(ctx, (final_kernel<MODE, CODE, false>), grid, kBalThreads, 0, src, st, npx, width, out, ctx->d_lab_gamma,
                  ctx->d_lab_cbrt);
    return BV_OK;
}

template <int MODE>
static int dispatch_final(bv_ctx *ctx, const uint8_t *src, const BalFrame *st, int batch, size_t npx, int width,
                          int code, const BalOutputs &out, bool vec) {
    switch (code) {
        case -1: return launch_final<MODE, -1>(ctx, src, st, batch, npx, width, out, vec);
        case BV_BGR2HSV: return launch_final<MODE, BV_BGR2HSV>(ctx, src, st, batch, npx, width, out, vec);
        case BV_BGR2LAB: return launch_final<MODE, BV_BGR2LAB>(ctx, src, st, batch, npx, width, out, vec);
        case BV_BGR2GRAY: return launch_final<MODE, BV_BGR2GRAY>(ctx, src, st, batch, npx, width, out, vec);
        case BV_BGR2YCRCB: return launch_final<MODE, BV_BGR2YCRCB>(ctx, src, st, batch, npx, width, out, vec);
        case BV_BGR2HLS: return launch_final<MODE, BV_BGR2HLS>(ctx, src, st, batch, npx, width, out, vec);
        default: set_error("stage: conversion code %d is not available in the fused pass", code); return BV_ERR_INVALID;
    }
}

static size_t l2_chunk_bytes() {
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("BV_L2_CHUNK_MB");
        long mb = e ? atol(e) : 40;
        if (mb < 1) mb = 1;
        v = (size_t)mb << 20;
    }
    return v;
}

static bool all_vec(const uint8_t *src, const BalOutputs &out, size_t npx, int batch) {
    return vec_ok(src, npx, batch, 3) && vec_ok(out.balanced, npx, batch, 3) && vec_ok(out.converted, npx, batch, 3) &&
           vec_ok(out.mask, npx, batch, 1);
}

int convert_run(bv_ctx *ctx, const uint8_t *src, int batch, int height, int width, int cvt_code, const BalOutputs &out) {
    const size_t npx = (size_t)height * width;
    const bool vec = all_vec(src, out, npx, batch);
    if (out.mask_bits && !(vec && width % 16 == 0)) {
        set_error("convert_run: bit-packed mask needs 16-byte aligned buffers and width %% 16 == 0");
        return BV_ERR_INVALID;
    }
    return dispatch_final<0>(ctx, src, nullptr, batch, npx, width, cvt_code, out, vec);
}

int balance_run(bv_ctx *ctx, const uint8_t *src, int batch, int height, int width, const bv_balance_params &prm,
                int cvt_code, const BalOutputs &out, bv_balance_stats *stats_host) {
    if (prm.hsi_contrast_correct) {
        set_error("colour balance: the HSI branch (color_balance.cpp:702-774) is not implemented");
        return BV_ERR_UNSUPPORTED;
    }
    if (prm.horizontal_blocks != 1 || prm.vertical_blocks != 1) {
        set_error("colour balance: tiled equalisation (horizontal/vertical_blocks > 1) is not implemented");
        return BV_ERR_UNSUPPORTED;
    }
    const size_t npx = (size_t)height * width;
    if (npx >= (1ull << 31)) {
        set_error("colour balance: frame too large for 32-bit histogram counters");
        return BV_ERR_INVALID;
    }
    const bool vec = all_vec(src, out, npx, batch);
    if (out.mask_bits && !(vec && width % 16 == 0)) {
        set_error("balance_run: bit-packed mask needs 16-byte aligned buffers and width %% 16 == 0");
        return BV_ERR_INVALID;
    }
    BV_TRY(ensure_scratch(ctx, SCR_BAL_STATE, sizeof(BalFrame) * (size_t)batch));
    BalFrame *st = (BalFrame *)ctx->scratch[SCR_BAL_STATE];
    BV_CUDA(cudaMemsetAsync(st, 0, sizeof(BalFrame) * (size_t)batch, ctx->stream));

    // chunk the batch so that one chunk's input stays in L2 across the three passes
    int chunk = (int)(l2_chunk_bytes() / (npx * 3));
    if (chunk < 1) chunk = 1;
    for (int f0 = 0; f0 < batch; f0 += chunk) {
        const int nf = batch - f0 < chunk ? batch - f0 : chunk;
        const uint8_t *csrc = src + (size_t)f0 * npx * 3;
        BalFrame *cst = st + f0;
        int bpf = (ctx->sm_count * 4 + nf - 1) / nf;
        const size_t need = (npx / 16 + kBalThreads - 1) / kBalThreads;
        if ((size_t)bpf > need) bpf = (int)(need ? need : 1);
        dim3 grid(bpf, nf);
        if (vec)
            BV_LAUNCH(ctx, hist_bgr_kernel<true>, grid, kBalThreads, 0, csrc, cst, npx);
        else
            BV_LAUNCH(ctx, hist_bgr_kernel<false>, grid, kBalThreads, 0, csrc, cst, npx);
        BV_LAUNCH(ctx, stats_bgr_kernel, nf, 256, 0, cst, npx, prm, ctx->d_pow_quarter);
        if (prm.hsv_contrast_correct) {
            if (vec)
                BV_LAUNCH(ctx, hist_sv_kernel<true>, grid, kBalThreads, 0, csrc, cst, npx);
            else
                BV_LAUNCH(ctx, hist_sv_kernel<false>, grid, kBalThreads, 0, csrc, cst, npx);
            BV_LAUNCH(ctx, stats_sv_kernel, nf, 256, 0, cst, npx);
        }
        BalOutputs co = out;
        if (co.balanced) co.balanced += (size_t)f0 * npx * 3;
        if (co.converted) co.converted += (size_t)f0 * npx * (cvt_code == BV_BGR2GRAY ? 1 : 3);
        if (co.mask) co.mask += (size_t)f0 * npx;
        if (co.mask_bits) co.mask_bits += (size_t)f0 * height * (((width + 31) / 32) * 2);
        if (prm.hsv_contrast_correct)
            BV_TRY(dispatch_final<2>(ctx, csrc, cst, nf, npx, width, cvt_code, co, vec));
        else
            BV_TRY(dispatch_final<1>(ctx, csrc, cst, nf, npx, width, cvt_code, co, vec));
    }
    if (stats_host) {
        BV_CUDA(cudaMemcpy2DAsync(stats_host, sizeof(bv_balance_stats), &st[0].stats, sizeof(BalFrame),
                                  sizeof(bv_balance_stats), batch, cudaMemcpyDeviceToHost, ctx->stream));
        BV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return BV_OK;
}

}  // namespace bv

using namespace bv;

extern "C" int bv_color_balance(bv_ctx *ctx, const uint8_t *src_dev, uint8_t *dst_dev, int batch, int height, int width,
                                const bv_balance_params *params, bv_balance_stats *stats_host) {
    BV_REQUIRE(ctx && src_dev && dst_dev && params, "null argument");
    BV_REQUIRE(batch > 0 && height > 0 && width > 0, "batch, height and width must be positive");
    BV_CUDA(cudaSetDevice(ctx->device));
    BalOutputs out;
    memset(&out, 0, sizeof(out));
    out.balanced = dst_dev;
    for (int k = 0; k < 3; ++k) out.hi[k] = 255;
    return balance_run(ctx, src_dev, batch, height, width, *params, -1, out, stats_host);
}
