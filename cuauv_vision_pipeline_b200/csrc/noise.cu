// Gaussian noise step of the preprocessor (modules/preprocessor.py:115-119):
//     noise = numpy.random.randn(*mat.shape) * sigma;  mat = clip(mat + noise, 0, 255).astype(uint8)
// numpy's global generator is the legacy RandomState: MT19937 words -> 53-bit doubles -> Marsaglia's polar method
// (numpy/random/src/legacy/legacy-distributions.c: legacy_gauss; .../mt19937/mt19937.h: mt19937_next_double), values
// filled in raster order, the second value of a pair kept for the next call.  The whole stream is replayed on the
// device from the generator's state, and the state after the call is handed back, so that a seeded run of a module
// produces the frames the reference produces and leaves numpy's generator where the reference leaves it.
//
//   mt_extend_kernel   the MT19937 recurrence x[k] = x[k-227] ^ twist(x[k-624], x[k-623]) is sequential with a
//                      dependency distance of 227 words: one CTA advances 227 words per barrier out of a shared-memory
//                      ring and streams the raw (untempered) words to global memory; a block of 624 raw words IS the
//                      generator state at that point.
//   polar_count_kernel one thread per polar attempt (4 words): accepted or not, counted per block.
//   polar_scan_kernel  exclusive scan of the block counts.
//   polar_apply_kernel accepted attempt number j yields values 2j and 2j+1 of the stream: each thread adds its two values
//                      to the pixels they belong to; the thread of the last pair reports how many attempts were used
//                      (-> position in the word stream) and the value left over for the next call.
#include "common.cuh"

namespace bv {

constexpr int kMtWords = 624, kMtLag = 227;
constexpr int kPolarThreads = 1024;

struct NoiseResult {
    unsigned long long accepted_total;  // accepted attempts among those generated
    unsigned long long attempts_used;   // attempts consumed up to and including the last pair needed
    double gauss;                       // second value of the last pair when it was not consumed
    int has_gauss;
    int pad;
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__global__ void __launch_bounds__(256) mt_extend_kernel(uint32_t *__restrict__ raw, unsigned n_total) {
    __shared__ uint32_t ring[1024];  // >= 624 + 227: a step reads [n-624, n) and writes [n, n+227)
    const unsigned t = threadIdx.x;
    for (unsigned i = t; i < kMtWords; i += 256) ring[i] = raw[i];
    __syncthreads();
    for (unsigned n = kMtWords; n < n_total; n += kMtLag) {
        const unsigned k = n + t;
        if (t < kMtLag && k < n_total) {
            const uint32_t v = ring[(k - kMtLag) & 1023] ^ mt_twist(ring[(k - kMtWords) & 1023], ring[(k - kMtWords + 1) & 1023]);
            ring[k & 1023] = v;
            raw[k] = v;
        }
        __syncthreads();
    }
}

// mt19937_next_double: (a * 2^26 + b) / 2^53 with a, b the top 27 / 26 bits of two words (all exact)
__device__ __forceinline__ double mt_double(uint32_t w0, uint32_t w1) {
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}

// One pass of legacy_gauss's do-while: x1, x2 uniform in [-1, 1), accepted inside the unit disc (origin excluded).
// Products and the sum are rounded separately (the reference is compiled without contraction).
__device__ __forceinline__ bool polar_attempt(const uint32_t *__restrict__ words, double &x1, double &x2, double &r2) {
    const uint32_t w0 = mt_temper(words[0]), w1 = mt_temper(words[1]), w2 = mt_temper(words[2]), w3 = mt_temper(words[3]);
    x1 = 2.0 * mt_double(w0, w1) - 1.0;
    x2 = 2.0 * mt_double(w2, w3) - 1.0;
    r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
    return r2 < 1.0 && r2 != 0.0;
}

__global__ void __launch_bounds__(kPolarThreads) polar_count_kernel(const uint32_t *__restrict__ stream, unsigned n_attempts,
                                                                   unsigned *__restrict__ counts) {
    const unsigned a = blockIdx.x * kPolarThreads + threadIdx.x;
    double x1, x2, r2;
    const bool ok = a < n_attempts && polar_attempt(stream + 4 * (size_t)a, x1, x2, r2);
    const int n = __syncthreads_count(ok);
    if (threadIdx.x == 0) counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(1024) polar_scan_kernel(const unsigned *__restrict__ counts, unsigned n_blocks,
                                                         unsigned long long *__restrict__ offsets, NoiseResult *res) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_blocks; base += 1024) {
        const unsigned i = base + t;
        const unsigned long long v = i < n_blocks ? counts[i] : 0;
        unsigned long long s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? warp_sums[warp - 1] : 0) + s - v;
        if (i < n_blocks) offsets[i] = before;
        __syncthreads();
        if (t == 1023) carry = before + v;
        __syncthreads();
    }
    if (t == 0) {
        res->accepted_total = carry;
        res->attempts_used = 0;
        res->gauss = 0.0;
        res->has_gauss = 0;
    }
}

__device__ __forceinline__ uint8_t add_noise_px(uint8_t px, double g, double sigma) {
    double v = __dadd_rn((double)px, __dmul_rn(g, sigma));  // mat + noise, float64
    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);            // numpy.clip(mat, 0., 255.)
    return (uint8_t)v;                                       // .astype(uint8): truncation
}

__global__ void __launch_bounds__(kPolarThreads) polar_apply_kernel(const uint32_t *__restrict__ stream, unsigned n_attempts,
                                                                   const unsigned long long *__restrict__ offsets,
                                                                   NoiseResult *res, const uint8_t *__restrict__ src,
                                                                   uint8_t *__restrict__ dst, unsigned long long n_values,
                                                                   unsigned long long n_pairs, int first_cached,
                                                                   double cached, double sigma) {
    __shared__ unsigned warp_sums[32];
    if (res->accepted_total < n_pairs) return;  // not enough attempts generated: the host extends the stream and retries
    const unsigned t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const unsigned a = blockIdx.x * kPolarThreads + t;
    if (a == 0 && first_cached) dst[0] = add_noise_px(src[0], cached, sigma);
    double x1 = 0, x2 = 0, r2 = 1;
    const bool ok = a < n_attempts && polar_attempt(stream + 4 * (size_t)a, x1, x2, r2);
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sums[lane], s = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        warp_sums[lane] = s - w;
    }
    __syncthreads();
    if (!ok) return;
    const unsigned long long j = offsets[blockIdx.x] + warp_sums[warp] + __popc(ballot & ((1u << lane) - 1u));
    if (j >= n_pairs) return;
    const double f = sqrt(-2.0 * log(r2) / r2);
    const double g0 = __dmul_rn(f, x2), g1 = __dmul_rn(f, x1);  // legacy_gauss returns f*x2 first and keeps f*x1
    const unsigned long long i0 = (unsigned long long)first_cached + 2 * j;
    dst[i0] = add_noise_px(src[i0], g0, sigma);
    if (i0 + 1 < n_values) {
        dst[i0 + 1] = add_noise_px(src[i0 + 1], g1, sigma);
    } else {
        res->gauss = g1;
        res->has_gauss = 1;
    }
    if (j == n_pairs - 1) res->attempts_used = (unsigned long long)a + 1;
}

}  // namespace bv

using namespace bv;

extern "C" int bv_add_gaussian_noise(bv_ctx *ctx, const uint8_t *src_dev, uint8_t *dst_dev, size_t n_values, double sigma,
                                     bv_mt19937_state *state) {
    BV_REQUIRE(ctx && src_dev && dst_dev && state, "null argument");
    BV_REQUIRE(n_values <= ((size_t)1 << 29), "too many values for one call");
    BV_REQUIRE(state->pos >= 0 && state->pos <= kMtWords, "generator position outside 0..624");
    if (n_values == 0) return BV_OK;
    BV_CUDA(cudaSetDevice(ctx->device));
    const int first_cached = state->has_gauss ? 1 : 0;
    const unsigned long long fresh = n_values - first_cached;
    const unsigned long long n_pairs = (fresh + 1) / 2;
    // acceptance probability pi/4; eight standard deviations of slack, extended and retried if it still falls short
    unsigned long long n_attempts = (unsigned long long)(n_pairs / 0.7853 + 8.0 * sqrt((double)n_pairs) + 64.0);
    if (n_pairs == 0) n_attempts = 0;
    NoiseResult res;
    memset(&res, 0, sizeof(res));
    unsigned n_total = 0;
    for (int round = 0;; ++round) {
        if (round == 4) {
            set_error("bv_add_gaussian_noise: the polar method rejected implausibly many attempts");
            return BV_ERR_CAPACITY;
        }
        const unsigned long long words = (unsigned long long)state->pos + 4 * n_attempts;
        const unsigned long long total = (words + kMtWords - 1) / kMtWords * kMtWords + kMtWords;
        BV_REQUIRE(total < (1ull << 32) && n_attempts < (1ull << 32) - kPolarThreads, "too many values for one call");
        n_total = (unsigned)total;
        const unsigned n_blocks = (unsigned)((n_attempts + kPolarThreads - 1) / kPolarThreads);
        BV_TRY(ensure_scratch(ctx, SCR_NOISE_RAW, sizeof(uint32_t) * (size_t)n_total));
        BV_TRY(ensure_scratch(ctx, SCR_NOISE_AUX, sizeof(NoiseResult) + (sizeof(unsigned long long) + sizeof(unsigned)) * ((size_t)n_blocks + 1)));
        uint32_t *raw = (uint32_t *)ctx->scratch[SCR_NOISE_RAW];
        NoiseResult *d_res = (NoiseResult *)ctx->scratch[SCR_NOISE_AUX];
        unsigned long long *offsets = (unsigned long long *)(d_res + 1);
        unsigned *counts = (unsigned *)(offsets + n_blocks + 1);
        BV_CUDA(cudaMemcpyAsync(raw, state->key, sizeof(uint32_t) * kMtWords, cudaMemcpyHostToDevice, ctx->stream));
        BV_LAUNCH(ctx, mt_extend_kernel, 1, 256, 0, raw, n_total);
        const uint32_t *stream = raw + state->pos;
        if (n_blocks) BV_LAUNCH(ctx, polar_count_kernel, n_blocks, kPolarThreads, 0, stream, (unsigned)n_attempts, counts);
        BV_LAUNCH(ctx, polar_scan_kernel, 1, 1024, 0, counts, n_blocks, offsets, d_res);
        if (n_blocks)
            BV_LAUNCH(ctx, polar_apply_kernel, n_blocks, kPolarThreads, 0, stream, (unsigned)n_attempts, offsets, d_res, src_dev,
                      dst_dev, (unsigned long long)n_values, n_pairs, first_cached, state->gauss, sigma);
        else  // a single value, taken from the cached one
            BV_LAUNCH(ctx, polar_apply_kernel, 1, kPolarThreads, 0, stream, 0u, offsets, d_res, src_dev, dst_dev,
                      (unsigned long long)n_values, n_pairs, first_cached, state->gauss, sigma);
        BV_CUDA(cudaMemcpyAsync(&res, d_res, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
        BV_CUDA(cudaStreamSynchronize(ctx->stream));
        if (res.accepted_total >= n_pairs) break;
        n_attempts = n_attempts + n_attempts / 8 + 4096;
    }
    // the generator after the call: raw block holding the last consumed word, position just behind it
    const unsigned long long q = (unsigned long long)state->pos + 4 * res.attempts_used;
    if (res.attempts_used) {
        const unsigned long long block = (q - 1) / kMtWords;
        BV_CUDA(cudaMemcpyAsync(state->key, (const uint32_t *)ctx->scratch[SCR_NOISE_RAW] + block * kMtWords,
                                sizeof(uint32_t) * kMtWords, cudaMemcpyDeviceToHost, ctx->stream));
        BV_CUDA(cudaStreamSynchronize(ctx->stream));
        state->pos = (int32_t)(q - block * kMtWords);
    }
    state->has_gauss = res.has_gauss;
    state->gauss = res.has_gauss ? res.gauss : 0.0;
    return BV_OK;
}
