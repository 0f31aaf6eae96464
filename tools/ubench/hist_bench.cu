// hist_bench.cu -- stand-alone micro-benchmarks that decide the design of the colour-balance
// passes (not part of the library; built by tools/ubench/build.sh, run under gpurun).
//
//   pass 1 variants (B,G,R histograms of 8 frames 2208x1242):
//     base      per-warp private 3x256 u32 histograms, ATOMS on random banks (the r01 kernel)
//     stripeS   one histogram per block whose counters are striped over S banks
//               (counter (v, lane % S) lives in bank (v*S + lane%S) % 32): S = 32 is conflict-free
//     stripe16p conflict-free, two 16-bit counters per word
//     floor     same loads and byte extraction, no atomics
//   pass 3 variants (mask bits of balance -> HSV -> inRange):
//     compute   table -> BGR2HSV -> S/V tables -> HSV2BGR -> BGR2HSV -> inRange  (exact arithmetic)
//     lut       table -> BGR2HSV -> S/V tables -> one bit gathered from a 1.47 MB table indexed by
//               (H, S', V') that the exact arithmetic filled once (frame independent)
//
// usage: hist_bench [frames.raw]   (raw uint8 BGR, 8 x 1242 x 2208 x 3; synthetic noise otherwise)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../cuauv_vision_pipeline_b200/csrc/pixel_math.cuh"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

static const int kH = 1242, kW = 2208, kNF = 8;
static const size_t kNpx = (size_t)kH * kW;

__device__ __forceinline__ uint4 ld_nc(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
struct Px16 {
    uint32_t w[12];
};
__device__ __forceinline__ void load16(const uint8_t *base, size_t g, Px16 &p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(base) + g * 3;
    const uint4 a = ld_nc(q), b = ld_nc(q + 1), c = ld_nc(q + 2);
    p.w[0] = a.x; p.w[1] = a.y; p.w[2] = a.z; p.w[3] = a.w;
    p.w[4] = b.x; p.w[5] = b.y; p.w[6] = b.z; p.w[7] = b.w;
    p.w[8] = c.x; p.w[9] = c.y; p.w[10] = c.z; p.w[11] = c.w;
}
#define GETB(W, k) (((W)[(k) >> 2] >> (8 * ((k)&3))) & 0xFFu)

// ------------------------------------------------------------------------------------------ pass 1
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) hist_base(const uint8_t *__restrict__ src, uint32_t *__restrict__ out, size_t npx) {
    __shared__ uint32_t h[WARPS][3][256];
    for (int i = threadIdx.x; i < WARPS * 768; i += blockDim.x) (&h[0][0][0])[i] = 0;
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t(*hw)[256] = h[threadIdx.x >> 5];
    const size_t stride = (size_t)gridDim.x * blockDim.x, ng = npx / 16;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
#pragma unroll
        for (int k = 0; k < 48; ++k) atomicAdd(&hw[k % 3][GETB(in.w, k)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += (&h[w][0][0])[i];
        if (s) atomicAdd(out + frame * 768 + i, s);
    }
}

template <int S>
__global__ void __launch_bounds__(1024) hist_stripe(const uint8_t *__restrict__ src, uint32_t *__restrict__ out, size_t npx) {
    extern __shared__ uint32_t cnt[];  // [3][256][S]
    for (int i = threadIdx.x; i < 768 * S; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t *mine = cnt + (threadIdx.x & (S - 1));
    const size_t stride = (size_t)gridDim.x * blockDim.x, ng = npx / 16;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
#pragma unroll
        for (int k = 0; k < 48; ++k) atomicAdd(mine + ((k % 3) * 256 + GETB(in.w, k)) * S, 1u);
    }
    __syncthreads();
    // 32 consecutive words = 32/S rows; segmented sum over S lanes
    for (int i = threadIdx.x; i < 768 * S; i += blockDim.x) {
        uint32_t v = cnt[i];
#pragma unroll
        for (int d = 1; d < S; d <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if ((i & (S - 1)) == 0 && v) atomicAdd(out + frame * 768 + i / S, v);
    }
}

// conflict-free, two 16-bit counters per word: cnt[3][128][32]
__global__ void __launch_bounds__(1024) hist_stripe16p(const uint8_t *__restrict__ src, uint32_t *__restrict__ out, size_t npx) {
    extern __shared__ uint32_t cnt[];
    for (int i = threadIdx.x; i < 3 * 128 * 32; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t *mine = cnt + (threadIdx.x & 31);
    const size_t stride = (size_t)gridDim.x * blockDim.x, ng = npx / 16;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
#pragma unroll
        for (int k = 0; k < 48; ++k) {
            const uint32_t v = GETB(in.w, k);
            atomicAdd(mine + ((k % 3) * 128 + (v >> 1)) * 32, (v & 1) ? 0x10000u : 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 128 * 32; i += blockDim.x) {
        const uint32_t v = cnt[i];
        const uint32_t lo = __reduce_add_sync(0xFFFFFFFFu, v & 0xFFFFu), hi = __reduce_add_sync(0xFFFFFFFFu, v >> 16);
        if ((i & 31) == 0) {
            const int row = i >> 5;  // c*128 + v/2
            const int c = row >> 7, v2 = (row & 127) * 2;
            if (lo) atomicAdd(out + frame * 768 + c * 256 + v2, lo);
            if (hi) atomicAdd(out + frame * 768 + c * 256 + v2 + 1, hi);
        }
    }
}

__global__ void __launch_bounds__(256) hist_floor(const uint8_t *__restrict__ src, uint32_t *__restrict__ out, size_t npx) {
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    const size_t stride = (size_t)gridDim.x * blockDim.x, ng = npx / 16;
    uint32_t acc[3] = {0, 0, 0};
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
#pragma unroll
        for (int k = 0; k < 48; ++k) acc[k % 3] += GETB(in.w, k);
    }
    if (acc[0] + acc[1] + acc[2] == 0xFFFFFFFFu) out[0] = 1;
}

// ------------------------------------------------------------------------------------------ pass 3
struct Tabs {
    uint8_t lut[3][256];
    uint8_t lut_sv[2][256];
};

__global__ void build_bits(uint32_t *__restrict__ bits, int lo0, int lo1, int lo2, int hi0, int hi1, int hi2) {
    __shared__ int sdiv[256], hdiv[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;  // (H*256 + S)*256 + V
    if (idx >= 180u * 65536u) return;
    const int H = idx >> 16, S = (idx >> 8) & 255, V = idx & 255;
    const uint32_t p = bv::hsv2bgr_packed(H, S, V, true);
    int h, s, v;
    bv::bgr2hsv((int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)(p >> 16), sdiv, hdiv, h, s, v);
    const bool in = h >= lo0 && h <= hi0 && s >= lo1 && s <= hi1 && v >= lo2 && v <= hi2;
    const uint32_t b = __ballot_sync(0xFFFFFFFFu, in);
    if ((threadIdx.x & 31) == 0) bits[idx >> 5] = b;
}

template <bool LUT>
__global__ void __launch_bounds__(256) final_mask(const uint8_t *__restrict__ src, const Tabs *__restrict__ tabs,
                                                  const uint32_t *__restrict__ bitlut, uint16_t *__restrict__ out, size_t npx,
                                                  int lo0, int lo1, int lo2, int hi0, int hi1, int hi2) {
    __shared__ Tabs t;
    __shared__ int sdiv[256], hdiv[256];
    for (int i = threadIdx.x; i < (int)sizeof(Tabs); i += blockDim.x) ((uint8_t *)&t)[i] = ((const uint8_t *)tabs)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    const uint32_t stride = gridDim.x * blockDim.x, ng = (uint32_t)(npx / 16);
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int h, s, v;
            bv::bgr2hsv(t.lut[0][GETB(in.w, 3 * j)], t.lut[1][GETB(in.w, 3 * j + 1)], t.lut[2][GETB(in.w, 3 * j + 2)], sdiv, hdiv,
                        h, s, v);
            s = t.lut_sv[0][s];
            v = t.lut_sv[1][v];
            if (LUT) {
                const uint32_t word = __ldg(bitlut + (((uint32_t)h * 256u + (uint32_t)s) * 8u + ((uint32_t)v >> 5)));
                bits |= ((word >> (v & 31)) & 1u) << j;
            } else {
                const uint32_t p = bv::hsv2bgr_packed(h, s, v, true);
                int h2, s2, v2;
                bv::bgr2hsv((int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)(p >> 16), sdiv, hdiv, h2, s2, v2);
                const bool in_r = h2 >= lo0 && h2 <= hi0 && s2 >= lo1 && s2 <= hi1 && v2 >= lo2 && v2 <= hi2;
                bits |= (in_r ? 1u : 0u) << j;
            }
        }
        out[(size_t)frame * ng + g] = (uint16_t)bits;
    }
}


// ------------------------------------------------------------------------------------------ pass 2
// STORE = false: the r01 pass 2 (S and V only).  STORE = true: full BGR2HSV, H,S,V written to an
// L2-resident scratch image so that pass 3 does not repeat table look-ups and the conversion.
__device__ __forceinline__ void st_na(uint4 *p, const uint4 &v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
template <bool STORE>
__global__ void __launch_bounds__(256) pass2(const uint8_t *__restrict__ src, const Tabs *__restrict__ tabs,
                                             uint32_t *__restrict__ out, uint8_t *__restrict__ hsv, size_t npx) {
    __shared__ uint32_t h[8][2][256];
    __shared__ uint8_t lut[3][256];
    __shared__ int sdiv[256], hdiv[256];
    for (int i = threadIdx.x; i < 8 * 512; i += blockDim.x) (&h[0][0][0])[i] = 0;
    for (int i = threadIdx.x; i < 768; i += blockDim.x) (&lut[0][0])[i] = (&tabs->lut[0][0])[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t(*hw)[256] = h[threadIdx.x >> 5];
    const size_t stride = (size_t)gridDim.x * blockDim.x, ng = npx / 16;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int hh, ss, vv;
            const int b = lut[0][GETB(in.w, 3 * j)], gg = lut[1][GETB(in.w, 3 * j + 1)], r = lut[2][GETB(in.w, 3 * j + 2)];
            if (STORE) {
                bv::bgr2hsv(b, gg, r, sdiv, hdiv, hh, ss, vv);
                const uint32_t p = (uint32_t)hh | ((uint32_t)ss << 8) | ((uint32_t)vv << 16);
                const int ofs = 3 * j, wi = ofs >> 2, sh = 8 * (ofs & 3);
                o[wi] |= p << sh;
                if (sh > 8) o[wi + 1] |= p >> (32 - sh);
            } else {
                vv = max(max(b, gg), r);
                const int diff = vv - min(min(b, gg), r);
                ss = (diff * sdiv[vv] + 2048) >> 12;
            }
            atomicAdd(&hw[0][ss], 1u);
            atomicAdd(&hw[1][vv], 1u);
        }
        if (STORE) {
            uint4 *q = reinterpret_cast<uint4 *>(hsv + (size_t)frame * npx * 3) + g * 3;
            st_na(q, make_uint4(o[0], o[1], o[2], o[3]));
            st_na(q + 1, make_uint4(o[4], o[5], o[6], o[7]));
            st_na(q + 2, make_uint4(o[8], o[9], o[10], o[11]));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (&h[w][0][0])[i];
        if (s) atomicAdd(out + frame * 512 + i, s);
    }
}

// ------------------------------------------------------------------------------------------ pass 3 from the HSV scratch
struct LabTabs {
    uint16_t gtab[256];
    uint16_t ctab[3072];
};
#include "../../cuauv_vision_pipeline_b200/csrc/lab_tables.inc"

// OUT 0: mask bits via the bit table; 1: LAB image via exact arithmetic
template <int OUT>
__global__ void __launch_bounds__(256) pass3_hsv(const uint8_t *__restrict__ hsv, const Tabs *__restrict__ tabs,
                                                 const uint32_t *__restrict__ bitlut, const LabTabs *__restrict__ lt,
                                                 uint16_t *__restrict__ out_bits, uint8_t *__restrict__ out_img, size_t npx) {
    __shared__ uint8_t lsv[2][256];
    __shared__ LabTabs lab;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&lsv[0][0])[i] = (&tabs->lut_sv[0][0])[i];
    if (OUT == 1)
        for (int i = threadIdx.x; i < (int)(sizeof(LabTabs) / 2); i += blockDim.x) ((uint16_t *)&lab)[i] = ((const uint16_t *)lt)[i];
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = hsv + (size_t)frame * npx * 3;
    const uint32_t stride = gridDim.x * blockDim.x, ng = (uint32_t)(npx / 16);
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
        uint32_t bits = 0;
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t h = GETB(in.w, 3 * j), s = lsv[0][GETB(in.w, 3 * j + 1)], v = lsv[1][GETB(in.w, 3 * j + 2)];
            if (OUT == 0) {
                const uint32_t word = __ldg(bitlut + ((h * 256u + s) * 8u + (v >> 5)));
                bits |= ((word >> (v & 31)) & 1u) << j;
            } else {
                const uint32_t p = bv::hsv2bgr_packed((int)h, (int)s, (int)v, true);
                int L, a, b;
                bv::bgr2lab((int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)(p >> 16), lab.gtab, lab.ctab, L, a, b);
                const uint32_t pc = (uint32_t)L | ((uint32_t)a << 8) | ((uint32_t)b << 16);
                const int ofs = 3 * j, wi = ofs >> 2, sh = 8 * (ofs & 3);
                o[wi] |= pc << sh;
                if (sh > 8) o[wi + 1] |= pc >> (32 - sh);
            }
        }
        if (OUT == 0) {
            out_bits[(size_t)frame * ng + g] = (uint16_t)bits;
        } else {
            uint4 *q = reinterpret_cast<uint4 *>(out_img + (size_t)frame * npx * 3) + g * 3;
            st_na(q, make_uint4(o[0], o[1], o[2], o[3]));
            st_na(q + 1, make_uint4(o[4], o[5], o[6], o[7]));
            st_na(q + 2, make_uint4(o[8], o[9], o[10], o[11]));
        }
    }
}

// the r01 pass 3 for C2: table -> BGR2HSV -> S/V tables -> HSV2BGR -> LAB
__global__ void __launch_bounds__(256) pass3_lab_bgr(const uint8_t *__restrict__ src, const Tabs *__restrict__ tabs,
                                                     const LabTabs *__restrict__ lt, uint8_t *__restrict__ out_img, size_t npx) {
    __shared__ Tabs t;
    __shared__ int sdiv[256], hdiv[256];
    __shared__ LabTabs lab;
    for (int i = threadIdx.x; i < (int)sizeof(Tabs); i += blockDim.x) ((uint8_t *)&t)[i] = ((const uint8_t *)tabs)[i];
    for (int i = threadIdx.x; i < (int)(sizeof(LabTabs) / 2); i += blockDim.x) ((uint16_t *)&lab)[i] = ((const uint16_t *)lt)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sdiv[i] = bv::hsv_sdiv(i);
        hdiv[i] = bv::hsv_hdiv(i);
    }
    __syncthreads();
    const int frame = blockIdx.y;
    const uint8_t *f = src + (size_t)frame * npx * 3;
    const uint32_t stride = gridDim.x * blockDim.x, ng = (uint32_t)(npx / 16);
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += stride) {
        Px16 in;
        load16(f, g, in);
        uint32_t o[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) o[k] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int h, s, v;
            bv::bgr2hsv(t.lut[0][GETB(in.w, 3 * j)], t.lut[1][GETB(in.w, 3 * j + 1)], t.lut[2][GETB(in.w, 3 * j + 2)], sdiv, hdiv,
                        h, s, v);
            const uint32_t p = bv::hsv2bgr_packed(h, t.lut_sv[0][s], t.lut_sv[1][v], true);
            int L, a, b;
            bv::bgr2lab((int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)(p >> 16), lab.gtab, lab.ctab, L, a, b);
            const uint32_t pc = (uint32_t)L | ((uint32_t)a << 8) | ((uint32_t)b << 16);
            const int ofs = 3 * j, wi = ofs >> 2, sh = 8 * (ofs & 3);
            o[wi] |= pc << sh;
            if (sh > 8) o[wi + 1] |= pc >> (32 - sh);
        }
        uint4 *q = reinterpret_cast<uint4 *>(out_img + (size_t)frame * npx * 3) + g * 3;
        st_na(q, make_uint4(o[0], o[1], o[2], o[3]));
        st_na(q + 1, make_uint4(o[4], o[5], o[6], o[7]));
        st_na(q + 2, make_uint4(o[8], o[9], o[10], o[11]));
    }
}

// ------------------------------------------------------------------------------------------ host
template <class F>
static float time_ms(F launch, int reps = 20) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char **argv) {
    const size_t bytes = kNpx * 3 * kNF;
    std::vector<uint8_t> host(bytes);
    bool loaded = false;
    if (argc > 1) {
        FILE *fp = fopen(argv[1], "rb");
        if (fp) {
            loaded = fread(host.data(), 1, bytes, fp) == bytes;
            fclose(fp);
        }
    }
    if (!loaded) {
        uint32_t s = 12345;
        const int mean[3] = {142, 107, 48};
        for (size_t i = 0; i < bytes; ++i) {
            int acc = 0;
            for (int k = 0; k < 4; ++k) {
                s = s * 1664525u + 1013904223u;
                acc += (s >> 24) & 31;
            }
            int v = mean[i % 3] + acc - 62;
            host[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    printf("input: %s\n", loaded ? argv[1] : "synthetic noise");
    uint8_t *d_src;
    uint32_t *d_hist, *d_ref;
    CK(cudaMalloc(&d_src, bytes));
    CK(cudaMemcpy(d_src, host.data(), bytes, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_hist, kNF * 768 * 4));
    CK(cudaMalloc(&d_ref, kNF * 768 * 4));
    // CPU reference histogram
    std::vector<uint32_t> ref(kNF * 768, 0);
    for (int fr = 0; fr < kNF; ++fr)
        for (size_t p = 0; p < kNpx; ++p)
            for (int c = 0; c < 3; ++c) ref[fr * 768 + c * 256 + host[((size_t)fr * kNpx + p) * 3 + c]]++;
    auto check = [&](const char *name) {
        std::vector<uint32_t> got(kNF * 768);
        CK(cudaMemcpy(got.data(), d_hist, got.size() * 4, cudaMemcpyDeviceToHost));
        // the timed loop ran 23 launches into the same buffer after a memset: counts are 23x
        bool ok = true;
        for (size_t i = 0; i < got.size(); ++i)
            if (got[i] != ref[i]) ok = false;
        printf("    %-28s %s\n", name, ok ? "histograms exact" : "MISMATCH");
    };
    auto run_hist = [&](const char *name, auto launch) {
        const float ms = time_ms(launch);
        CK(cudaMemset(d_hist, 0, kNF * 768 * 4));
        launch();
        CK(cudaDeviceSynchronize());
        printf("%-34s %8.2f us/frame  (%6.1f GB/s)\n", name, ms * 1e3 / kNF, bytes / ms / 1e6);
        check(name);
    };
    CK(cudaFuncSetAttribute(hist_stripe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 768 * 32 * 4));
    CK(cudaFuncSetAttribute(hist_stripe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 768 * 16 * 4));
    CK(cudaFuncSetAttribute(hist_stripe16p, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 128 * 32 * 4));
    if (argc > 2 && !strcmp(argv[2], "ncu")) {  // one launch of each kernel of interest, for ncu --set full
        Tabs tb;
        for (int i = 0; i < 256; ++i) {
            tb.lut[0][i] = (uint8_t)i;
            tb.lut[1][i] = (uint8_t)(i * 4 / 3 > 255 ? 255 : i * 4 / 3);
            tb.lut[2][i] = (uint8_t)(i * 29 / 10 > 255 ? 255 : i * 29 / 10);
            int x = i < 4 ? 4 : i > 173 ? 173 : i;
            tb.lut_sv[0][i] = (uint8_t)((x - 4) * 255 / 169);
            x = i < 111 ? 111 : i > 235 ? 235 : i;
            tb.lut_sv[1][i] = (uint8_t)((x - 111) * 255 / 124);
        }
        Tabs *dt;
        uint32_t *db, *dh2;
        uint16_t *dm;
        uint8_t *dhsv, *dimg;
        LabTabs lt, *dlt;
        memcpy(lt.gtab, kLabGammaTab, sizeof lt.gtab);
        memcpy(lt.ctab, kLabCbrtTab, sizeof lt.ctab);
        CK(cudaMalloc(&dlt, sizeof lt));
        CK(cudaMemcpy(dlt, &lt, sizeof lt, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&dt, sizeof tb));
        CK(cudaMemcpy(dt, &tb, sizeof tb, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&db, 180 * 65536 / 8));
        CK(cudaMalloc(&dm, kNF * kNpx / 16 * 2));
        CK(cudaMalloc(&dhsv, bytes));
        CK(cudaMalloc(&dimg, bytes));
        CK(cudaMalloc(&dh2, kNF * 512 * 4));
        CK(cudaMemset(dh2, 0, kNF * 512 * 4));
        build_bits<<<180 * 65536 / 256, 256>>>(db, 10, 20, 60, 30, 100, 255);
        hist_base<8><<<dim3(74, kNF), 256>>>(d_src, d_hist, kNpx);
        pass2<false><<<dim3(74, kNF), 256>>>(d_src, dt, dh2, dhsv, kNpx);
        pass2<true><<<dim3(74, kNF), 256>>>(d_src, dt, dh2, dhsv, kNpx);
        pass3_hsv<0><<<dim3(74, kNF), 256>>>(dhsv, dt, db, dlt, dm, dimg, kNpx);
        pass3_hsv<1><<<dim3(74, kNF), 256>>>(dhsv, dt, db, dlt, dm, dimg, kNpx);
        pass3_lab_bgr<<<dim3(74, kNF), 256>>>(d_src, dt, dlt, dimg, kNpx);
        final_mask<true><<<dim3(74, kNF), 256>>>(d_src, dt, db, dm, kNpx, 10, 20, 60, 30, 100, 255);
        final_mask<false><<<dim3(74, kNF), 256>>>(d_src, dt, db, dm, kNpx, 10, 20, 60, 30, 100, 255);
        CK(cudaDeviceSynchronize());
        printf("ncu pass done\n");
        return 0;
    }
    for (int bpf : {74, 148, 296, 592}) {
        char nm[64];
        snprintf(nm, sizeof nm, "base<8 warps> bpf=%d", bpf);
        run_hist(nm, [&] { hist_base<8><<<dim3(bpf, kNF), 256>>>(d_src, d_hist, kNpx); });
    }
    for (int bpf : {74, 148, 296}) {
        char nm[64];
        snprintf(nm, sizeof nm, "base<16 warps> bpf=%d", bpf);
        run_hist(nm, [&] { hist_base<16><<<dim3(bpf, kNF), 512>>>(d_src, d_hist, kNpx); });
    }
    for (int thr : {512, 1024})
        for (int bpf : {19, 37, 74}) {
            char nm[64];
            snprintf(nm, sizeof nm, "stripe32 thr=%d bpf=%d", thr, bpf);
            run_hist(nm, [&] { hist_stripe<32><<<dim3(bpf, kNF), thr, 768 * 32 * 4>>>(d_src, d_hist, kNpx); });
            snprintf(nm, sizeof nm, "stripe16 thr=%d bpf=%d", thr, bpf);
            run_hist(nm, [&] { hist_stripe<16><<<dim3(bpf, kNF), thr, 768 * 16 * 4>>>(d_src, d_hist, kNpx); });
            snprintf(nm, sizeof nm, "stripe8 thr=%d bpf=%d", thr, bpf);
            run_hist(nm, [&] { hist_stripe<8><<<dim3(bpf, kNF), thr, 768 * 8 * 4>>>(d_src, d_hist, kNpx); });
            snprintf(nm, sizeof nm, "stripe16p thr=%d bpf=%d", thr, bpf);
            run_hist(nm, [&] { hist_stripe16p<<<dim3(bpf, kNF), thr, 3 * 128 * 32 * 4>>>(d_src, d_hist, kNpx); });
        }
    for (int bpf : {148, 592}) {
        char nm[64];
        snprintf(nm, sizeof nm, "floor bpf=%d", bpf);
        const float ms = time_ms([&] { hist_floor<<<dim3(bpf, kNF), 256>>>(d_src, d_hist, kNpx); });
        printf("%-34s %8.2f us/frame  (%6.1f GB/s)\n", nm, ms * 1e3 / kNF, bytes / ms / 1e6);
    }

    // ---- pass 3 ----
    Tabs tabs;
    for (int i = 0; i < 256; ++i) {
        auto clipgain = [&](int lo, int hi, double g) {
            int x = i < lo ? lo : i > hi ? hi : i;
            double v = x * g;
            return (uint8_t)(v > 255 ? 255 : (int)v);
        };
        tabs.lut[0][i] = clipgain(60, 230, 1.0);
        tabs.lut[1][i] = clipgain(40, 190, 1.33);
        tabs.lut[2][i] = clipgain(8, 120, 2.9);
        auto stretch = [&](int lo, int hi) {
            int x = i < lo ? lo : i > hi ? hi : i;
            return (uint8_t)(((x - lo) * 255) / (hi - lo));
        };
        tabs.lut_sv[0][i] = stretch(4, 173);
        tabs.lut_sv[1][i] = stretch(111, 235);
    }
    Tabs *d_tabs;
    uint32_t *d_bits;
    uint16_t *d_m0, *d_m1;
    CK(cudaMalloc(&d_tabs, sizeof(Tabs)));
    CK(cudaMemcpy(d_tabs, &tabs, sizeof(Tabs), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bits, 180 * 65536 / 8));
    CK(cudaMalloc(&d_m0, kNF * kNpx / 16 * 2));
    CK(cudaMalloc(&d_m1, kNF * kNpx / 16 * 2));
    const int los[2][3] = {{10, 20, 60}, {0, 40, 60}}, his[2][3] = {{30, 100, 255}, {179, 255, 255}};
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int *lo = los[cfg], *hi = his[cfg];
        const float msb = time_ms([&] { build_bits<<<180 * 65536 / 256, 256>>>(d_bits, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]); }, 5);
        printf("bounds cfg %d: build_bits %.1f us\n", cfg, msb * 1e3);
        for (int bps : {4, 8}) {
            const int bpf = 148 * bps / kNF > 0 ? 148 * bps / kNF : 1;
            const float m0 = time_ms([&] {
                final_mask<false><<<dim3(bpf, kNF), 256>>>(d_src, d_tabs, d_bits, d_m0, kNpx, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
            });
            const float m1 = time_ms([&] {
                final_mask<true><<<dim3(bpf, kNF), 256>>>(d_src, d_tabs, d_bits, d_m1, kNpx, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
            });
            std::vector<uint16_t> a(kNF * kNpx / 16), b(kNF * kNpx / 16);
            CK(cudaMemcpy(a.data(), d_m0, a.size() * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(b.data(), d_m1, b.size() * 2, cudaMemcpyDeviceToHost));
            size_t set = 0;
            for (auto w : a) set += __builtin_popcount(w);
            printf("  final_mask bpf=%d: compute %.2f us/frame, lut %.2f us/frame, masks %s (%.2f %% set)\n", bpf, m0 * 1e3 / kNF,
                   m1 * 1e3 / kNF, a == b ? "identical" : "DIFFER", 100.0 * set / (kNF * kNpx));
        }
    }
    // ---- pass 2 and pass 3 over the HSV scratch (4-frame chunks, as the library would run them) ----
    {
        uint8_t *d_hsv, *d_img0, *d_img1;
        uint32_t *d_h2;
        LabTabs lt;
        memcpy(lt.gtab, kLabGammaTab, sizeof lt.gtab);
        memcpy(lt.ctab, kLabCbrtTab, sizeof lt.ctab);
        LabTabs *d_lt;
        CK(cudaMalloc(&d_lt, sizeof lt));
        CK(cudaMemcpy(d_lt, &lt, sizeof lt, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_hsv, bytes));
        CK(cudaMalloc(&d_img0, bytes));
        CK(cudaMalloc(&d_img1, bytes));
        CK(cudaMalloc(&d_h2, kNF * 512 * 4));
        CK(cudaMemset(d_h2, 0, kNF * 512 * 4));
        const int lo[3] = {10, 20, 60}, hi[3] = {30, 100, 255};
        build_bits<<<180 * 65536 / 256, 256>>>(d_bits, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
        for (int nf : {4, 8})
            for (int bps : {4, 8}) {
                const int bpf = 148 * bps / nf;
                const float a = time_ms([&] { pass2<false><<<dim3(bpf, nf), 256>>>(d_src, d_tabs, d_h2, d_hsv, kNpx); });
                const float b = time_ms([&] { pass2<true><<<dim3(bpf, nf), 256>>>(d_src, d_tabs, d_h2, d_hsv, kNpx); });
                const float c = time_ms([&] { pass3_hsv<0><<<dim3(bpf, nf), 256>>>(d_hsv, d_tabs, d_bits, d_lt, d_m1, d_img1, kNpx); });
                const float d = time_ms([&] { pass3_hsv<1><<<dim3(bpf, nf), 256>>>(d_hsv, d_tabs, d_bits, d_lt, d_m1, d_img1, kNpx); });
                const float e = time_ms([&] { pass3_lab_bgr<<<dim3(bpf, nf), 256>>>(d_src, d_tabs, d_lt, d_img0, kNpx); });
                printf("nf=%d bpf=%d  us/frame: pass2 S,V only %.2f | pass2 + HSV store %.2f | pass3 bits from HSV %.2f | pass3 LAB from HSV %.2f | "
                       "pass3 LAB from BGR (r01) %.2f\n", nf, bpf, a * 1e3 / nf, b * 1e3 / nf, c * 1e3 / nf, d * 1e3 / nf, e * 1e3 / nf);
            }
        // correctness of the split: bits and LAB through the scratch == direct
        final_mask<false><<<dim3(74, kNF), 256>>>(d_src, d_tabs, d_bits, d_m0, kNpx, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
        pass2<true><<<dim3(74, kNF), 256>>>(d_src, d_tabs, d_h2, d_hsv, kNpx);
        pass3_hsv<0><<<dim3(74, kNF), 256>>>(d_hsv, d_tabs, d_bits, d_lt, d_m1, d_img1, kNpx);
        std::vector<uint16_t> a(kNF * kNpx / 16), b(kNF * kNpx / 16);
        CK(cudaMemcpy(a.data(), d_m0, a.size() * 2, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_m1, b.size() * 2, cudaMemcpyDeviceToHost));
        printf("mask bits via scratch + table vs direct arithmetic: %s\n", a == b ? "identical" : "DIFFER");
        pass3_hsv<1><<<dim3(74, kNF), 256>>>(d_hsv, d_tabs, d_bits, d_lt, d_m1, d_img1, kNpx);
        pass3_lab_bgr<<<dim3(74, kNF), 256>>>(d_src, d_tabs, d_lt, d_img0, kNpx);
        std::vector<uint8_t> i0(bytes), i1(bytes);
        CK(cudaMemcpy(i0.data(), d_img0, bytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(i1.data(), d_img1, bytes, cudaMemcpyDeviceToHost));
        printf("LAB via scratch vs direct: %s\n", i0 == i1 ? "identical" : "DIFFER");
    }
    return 0;
}
