// hsi.cu -- the HSI contrast-correction branch of process_frame (color_balance.cpp:702-774 with
// conv_rgb_to_hsi 167-261, conv_hsi_to_rgb 263-341, percentile_min_max_qselect 144-154), P2.
//
// It runs after the RGB / HSV stages on the balanced BGR frame:
//   forward   I = (r+g+b)/3, S = 1 - min/I, H = acos((r - g/2 - b/2) / sqrt(r^2+g^2+b^2-rg-rb-gb)),
//             2 pi - H when b > g; float32 planes, clipped to [0,2pi], [0,1], [0,255] (NaN -> lower bound)
//   bounds    the 0.2 % / 99.8 % order statistics of S and of I (the reference's quickselect; its
//             std::rand() pivots do not change the value it returns) -- four radix selects at once on the
//             device, no host round trip
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

// ---- the four order statistics (S and I, 0.2 % and 99.8 %) without leaving the device: radix select, 8 bits
// per round; a round is one histogram pass over both planes (four histograms: each plane against its two
// current prefixes) and a one-block step that narrows the four prefixes ----
struct Select4 {
    uint32_t prefix[4];  // S low, S high, I low, I high
    uint32_t rank[4];    // rank of the wanted element among those matching the prefix
    uint32_t hist[4][256];
};

__device__ __forceinline__ uint32_t float_key(float f) {  // order-preserving integer image of a float
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void select4_init_kernel(Select4 *st, uint32_t lo_k, uint32_t hi_k) {
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
    if (threadIdx.x < 4) {
        st->prefix[threadIdx.x] = 0;
        st->rank[threadIdx.x] = (threadIdx.x & 1) ? hi_k : lo_k;
    }
}

__global__ void __launch_bounds__(256) select4_hist_kernel(const float *__restrict__ S, const float *__restrict__ I, size_t n,
                                                           Select4 *__restrict__ st, uint32_t mask, int shift) {
    __shared__ uint32_t h[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1], p2 = st->prefix[2], p3 = st->prefix[3];
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t ks = float_key(S[i]), ki = float_key(I[i]);
        const uint32_t bs = (ks >> shift) & 0xFFu, bi = (ki >> shift) & 0xFFu;
        if ((ks & mask) == p0) atomicAdd(&h[0][bs], 1u);
        if ((ks & mask) == p1) atomicAdd(&h[1][bs], 1u);
        if ((ki & mask) == p2) atomicAdd(&h[2][bi], 1u);
        if ((ki & mask) == p3) atomicAdd(&h[3][bi], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)
        if ((&h[0][0])[i]) atomicAdd(&st->hist[0][0] + i, (&h[0][0])[i]);
}

// one block of 4 warps: warp j walks histogram j to the bin that holds rank[j]
__global__ void select4_step_kernel(Select4 *st, int shift) {
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = st->rank[j];
    uint32_t before = 0;
    int found = -1;
    for (int base = 0; base < 256 && found < 0; base += 32) {
        const uint32_t c = st->hist[j][base + lane];
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t hit = __ballot_sync(0xFFFFFFFFu, rank < before + incl);
        if (hit) {
            const int l = __ffs(hit) - 1;
            const uint32_t excl = __shfl_sync(0xFFFFFFFFu, incl - c, l);
            found = base + l;
            rank -= before + excl;
        } else {
            before += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    __syncwarp();
    for (int b = lane; b < 256; b += 32) st->hist[j][b] = 0;  // ready for the next round
    if (lane == 0) {
        st->prefix[j] |= (uint32_t)(found < 0 ? 255 : found) << shift;
        st->rank[j] = rank;
    }
}

__global__ void hsi_bounds_kernel(const Select4 *st, HsiBounds *bd) {
    float v[4];
    for (int j = 0; j < 4; ++j) {
        const uint32_t k = st->prefix[j];
        v[j] = __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
    }
    bd->s_min = v[0];
    bd->s_max = v[1];
    bd->i_min = v[2];
    bd->i_max = v[3];
    bd->s_mult = (float)(1. / (double)__fsub_rn(v[1], v[0]));      // 750-751: float subtraction, double division
    bd->i_mult = (float)(255. / (double)__fsub_rn(v[3], v[2]));
}

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
                                                           const HsiBounds *__restrict__ bounds) {
    const HsiBounds bd = *bounds;
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

// In-place HSI contrast correction of `batch` balanced BGR frames; everything stays on the stream (no host round trip).
int hsi_run(bv_ctx *ctx, uint8_t *bgr, int batch, size_t npx) {
    BV_TRY(ensure_scratch(ctx, SCR_HSI, sizeof(float) * 3 * npx + sizeof(Select4) + sizeof(HsiBounds) + 64));
    float *H = (float *)ctx->scratch[SCR_HSI], *S = H + npx, *I = S + npx;
    Select4 *sel = (Select4 *)(I + npx);
    HsiBounds *bd = (HsiBounds *)(sel + 1);
    const int grid = grid_for(ctx, npx, 256, 8);
    size_t lo_k = (size_t)(int)(0.002f * (float)npx), hi_k = (size_t)(int)(0.998f * (float)npx);  // 145-146
    if (lo_k >= npx) lo_k = npx - 1;
    if (hi_k >= npx) hi_k = npx - 1;
    for (int f = 0; f < batch; ++f) {
        uint8_t *frame = bgr + (size_t)f * npx * 3;
        BV_LAUNCH(ctx, hsi_forward_kernel, grid, 256, 0, frame, H, S, I, npx);
        BV_LAUNCH(ctx, select4_init_kernel, 1, 256, 0, sel, (uint32_t)lo_k, (uint32_t)hi_k);
        uint32_t mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            BV_LAUNCH(ctx, select4_hist_kernel, grid, 256, 0, S, I, npx, sel, mask, shift);
            BV_LAUNCH(ctx, select4_step_kernel, 1, 128, 0, sel, shift);
            mask |= 0xFFu << shift;
        }
        BV_LAUNCH(ctx, hsi_bounds_kernel, 1, 1, 0, sel, bd);
        BV_LAUNCH(ctx, hsi_backward_kernel, grid, 256, 0, H, S, I, frame, npx, bd);
    }
    return BV_OK;
}

}  // namespace bv
