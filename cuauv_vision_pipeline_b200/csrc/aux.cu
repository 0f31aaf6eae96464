// aux.cu -- the ZED auxiliary-plane conversions: pure streaming kernels.
//
//   rgba_to_rgb        capture_sources/zed.cpp:54-71, capture_sources/zed.py:49-50 (cv2 RGBA2RGB):
//                      drop the alpha byte, 4 B/px in -> 3 B/px out
//   normals_to_rgb01   capture_sources/zed.cpp:73-91: float4 normals -> 3 x float (v + 1) * 0.5
//   f32 -> u8 display  modules/poster.py:41-47, modules/record.py:106-113, modules/calibrate.py:111:
//                      clip((x - sub) / div * 255, 0, 255).astype(uint8)  and the clip-then-scale form
//   channel means      modules/auto_calibrate_zed.py:82 np.mean(img, axis=(0, 1)): exact integer sums
//
// Float32 arithmetic follows numpy operation by operation (python scalars are weak: they are
// converted to float32 first); the final cast truncates like astype(np.uint8).  NaN / inf inputs
// (invalid depth) are undefined in the numpy cast; here NaN -> 0, +inf -> 255, -inf -> 0.
#include "common.cuh"

namespace bv {

// 16 pixels per thread: 4 x 128-bit loads (64 B), 3 x 128-bit stores (48 B)
// SWAP: also exchange bytes 0 and 2 of every pixel (RGBA -> BGR), one byte permute per pixel
template <bool SWAP>
__global__ void __launch_bounds__(256) rgba_to_rgb_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                          size_t npx, bool vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t ngroups = vec ? npx / 16 : 0;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const uint4 *q = reinterpret_cast<const uint4 *>(src) + g * 4;
        const uint4 a = ld_stream(q), b = ld_stream(q + 1), c = ld_stream(q + 2), d = ld_stream(q + 3);
        const uint32_t in[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 RGBA words -> 3 packed words
            const uint32_t sel = SWAP ? 0x4012u : 0x4210u;  // byte 3 <- 0
            const uint32_t p0 = __byte_perm(in[4 * k], 0u, sel), p1 = __byte_perm(in[4 * k + 1], 0u, sel),
                           p2 = __byte_perm(in[4 * k + 2], 0u, sel), p3 = __byte_perm(in[4 * k + 3], 0u, sel);
            o[3 * k] = p0 | (p1 << 24);
            o[3 * k + 1] = (p1 >> 8) | (p2 << 16);
            o[3 * k + 2] = (p2 >> 16) | (p3 << 8);
        }
        uint4 *w = reinterpret_cast<uint4 *>(dst) + g * 3;
        st_stream(w, make_uint4(o[0], o[1], o[2], o[3]));
        st_stream(w + 1, make_uint4(o[4], o[5], o[6], o[7]));
        st_stream(w + 2, make_uint4(o[8], o[9], o[10], o[11]));
    }
    for (size_t p = ngroups * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
        dst[3 * p] = src[4 * p + (SWAP ? 2 : 0)];
        dst[3 * p + 1] = src[4 * p + 1];
        dst[3 * p + 2] = src[4 * p + (SWAP ? 0 : 2)];
    }
}

__global__ void __launch_bounds__(256) normals_to_rgb01_kernel(const float4 *__restrict__ src, float *__restrict__ dst, size_t npx) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
        const float4 v = __ldg(src + p);
        dst[3 * p] = __fmul_rn(__fadd_rn(v.x, 1.f), 0.5f);
        dst[3 * p + 1] = __fmul_rn(__fadd_rn(v.y, 1.f), 0.5f);
        dst[3 * p + 2] = __fmul_rn(__fadd_rn(v.z, 1.f), 0.5f);
    }
}

__device__ __forceinline__ uint32_t f32_to_u8_trunc(float y) {
    if (!(y == y)) return 0u;  // NaN
    y = fminf(fmaxf(y, 0.f), 255.f);
    return (uint32_t)__float2int_rz(y);
}

// mode 0: clip(((x - sub) / div) * 255, 0, 255)            (modules/poster.py:41-47; sub = 0, div = 1 for normals)
// mode 1: (clip((x - sub) / div, 0, 1) * 255)               (modules/record.py:106-109)
__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float *__restrict__ src, uint8_t *__restrict__ dst, size_t n,
                                                        float sub, float div, int mode, bool apply_affine, bool vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t ngroups = vec ? n / 16 : 0;
    auto one = [&](float x) -> uint32_t {
        if (apply_affine) x = __fdiv_rn(__fsub_rn(x, sub), div);
        if (mode == 1) {
            if (!(x == x)) return 0u;
            x = fminf(fmaxf(x, 0.f), 1.f);
            return (uint32_t)__float2int_rz(__fmul_rn(x, 255.f));
        }
        return f32_to_u8_trunc(__fmul_rn(x, 255.f));
    };
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
        const uint4 *q = reinterpret_cast<const uint4 *>(src) + g * 4;
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = ld_stream(q + k);
            o[k] = one(__uint_as_float(v.x)) | (one(__uint_as_float(v.y)) << 8) | (one(__uint_as_float(v.z)) << 16) |
                   (one(__uint_as_float(v.w)) << 24);
        }
        st_stream(reinterpret_cast<uint4 *>(dst) + g, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (size_t i = ngroups * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = (uint8_t)one(src[i]);
}

// exact per-channel sums of an interleaved uint8 image (np.mean(img, axis=(0,1)) numerators)
__global__ void __launch_bounds__(256) channel_sums_kernel(const uint8_t *__restrict__ src, size_t npx, int cn,
                                                           unsigned long long *__restrict__ sums) {
    __shared__ unsigned long long warp_sums[8][4];
    unsigned long long acc[4] = {0, 0, 0, 0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride)
        for (int c = 0; c < cn; ++c) acc[c] += src[p * cn + c];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = 0; c < cn; ++c) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[c] += __shfl_down_sync(0xFFFFFFFFu, acc[c], d);
        if (lane == 0) warp_sums[wid][c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < cn) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sums[w][threadIdx.x];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

}  // namespace bv

namespace bv {
int rgba_to_rgb_run(bv_ctx *ctx, const uint8_t *src_dev, uint8_t *dst_dev, size_t n_pixels, int swap_rb) {
    const bool vec = host_aligned16(src_dev) && host_aligned16(dst_dev);
    const int grid = grid_for(ctx, vec ? (n_pixels + 15) / 16 : n_pixels, 256, 8);
    if (swap_rb)
        BV_LAUNCH(ctx, rgba_to_rgb_kernel<true>, grid, 256, 0, src_dev, dst_dev, n_pixels, vec);
    else
        BV_LAUNCH(ctx, rgba_to_rgb_kernel<false>, grid, 256, 0, src_dev, dst_dev, n_pixels, vec);
    return BV_OK;
}
}  // namespace bv

using namespace bv;

extern "C" int bv_rgba_to_rgb(bv_ctx *ctx, const uint8_t *src_dev, uint8_t *dst_dev, size_t n_pixels) {
    BV_REQUIRE(ctx && src_dev && dst_dev, "null argument");
    BV_CUDA(cudaSetDevice(ctx->device));
    if (!n_pixels) return BV_OK;
    return rgba_to_rgb_run(ctx, src_dev, dst_dev, n_pixels, 0);
}

extern "C" int bv_normals_to_rgb01(bv_ctx *ctx, const float *src_xyzw_dev, float *dst_rgb_dev, size_t n_pixels) {
    BV_REQUIRE(ctx && src_xyzw_dev && dst_rgb_dev, "null argument");
    BV_REQUIRE(host_aligned16(src_xyzw_dev), "float4 source must be 16-byte aligned");
    BV_CUDA(cudaSetDevice(ctx->device));
    if (!n_pixels) return BV_OK;
    BV_LAUNCH(ctx, normals_to_rgb01_kernel, grid_for(ctx, n_pixels, 256, 8), 256, 0, (const float4 *)src_xyzw_dev, dst_rgb_dev,
              n_pixels);
    return BV_OK;
}

extern "C" int bv_f32_to_u8(bv_ctx *ctx, const float *src_dev, uint8_t *dst_dev, size_t n, int apply_affine, float sub,
                            float div, int clip_before_scale) {
    BV_REQUIRE(ctx && src_dev && dst_dev, "null argument");
    BV_CUDA(cudaSetDevice(ctx->device));
    if (!n) return BV_OK;
    const bool vec = host_aligned16(src_dev) && host_aligned16(dst_dev);
    BV_LAUNCH(ctx, f32_to_u8_kernel, grid_for(ctx, vec ? (n + 15) / 16 : n, 256, 8), 256, 0, src_dev, dst_dev, n, sub, div,
              clip_before_scale ? 1 : 0, apply_affine != 0, vec);
    return BV_OK;
}

extern "C" int bv_channel_sums(bv_ctx *ctx, const uint8_t *src_dev, size_t n_pixels, int channels, uint64_t *sums_dev) {
    BV_REQUIRE(ctx && src_dev && sums_dev, "null argument");
    BV_REQUIRE(channels >= 1 && channels <= 4, "channels must be 1..4");
    BV_CUDA(cudaSetDevice(ctx->device));
    BV_CUDA(cudaMemsetAsync(sums_dev, 0, sizeof(uint64_t) * channels, ctx->stream));
    if (!n_pixels) return BV_OK;
    BV_LAUNCH(ctx, channel_sums_kernel, grid_for(ctx, n_pixels, 256, 4), 256, 0, src_dev, n_pixels, channels,
              (unsigned long long *)sums_dev);
    return BV_OK;
}
