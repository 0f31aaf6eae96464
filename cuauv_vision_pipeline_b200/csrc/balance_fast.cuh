// balance_fast.cuh -- passes 2 and 3 of the colour balance for the common case (default tiling, 16-byte aligned frames of
// whole 16-pixel groups) with every per-pixel table look-up free of shared-memory bank conflicts.  Included by balance.cu
// after its statistics helpers; the generic kernels there remain the path for odd shapes and tiled equalisation.
//
// r01 measured the three passes on the shared-memory pipe, not on HBM or the issue slots (profiles/r01_ncu_c2_full_v3.md:
// 6.4 M wavefronts for 2.8 M look-up instructions in pass 3, 55 % of them conflicts): a 256-entry byte table occupies two
// words per bank, a warp's 32 random indices collide 1.5-2.7 times per instruction.  Here each table is replicated once per
// LANE: a "row" of 256 bytes per index value holds 32 lanes x 2 words, lane l only ever touches bank l, so every look-up is
// one wavefront.  The row stride of 256 bytes makes the address of a look-up a single byte permute of the packed input word:
//   PRMT(word, lane*4, 0x55k4) = (byte k of word) << 8 | lane*4        (lane*4 < 128 fits the low byte)
// A 64 KB region therefore carries two 32-bit tables (words 0 and 1 of every lane slot), which the passes use as
//   pass 2:  word 0 = (LUT_b[x], LUT_g[x], LUT_r[x], 0)   word 1 = sdiv[x]      (x = raw byte / V)
//   pass 3:  word 0 = float(2^23 + S'[x])                 word 1 = float(2^23 + V'[x])
// where the pass-3 entries are already the biased floats HSV -> BGR starts from (pixel_math.cuh), which also removes the
// two integer -> float steps per pixel.  hdiv (pass 2) and the Lab gamma table (pass 3) are indexed by computed values and
// live in 8-fold (hdiv: bank = 8 (i mod 4) + lane mod 8) / 32-fold (gamma: stride 128) replicated arrays.
#pragma once

namespace bv {

constexpr int kFastThreads = 512;
constexpr int kFastWarps = kFastThreads / 32;
constexpr uint32_t kRepBytes = 65536;  // 256 rows x (32 lanes x 2 words)

__device__ __forceinline__ uint32_t rep_addr(uint32_t word, uint32_t lane4, int k) {
    // (byte k of word) << 8 | lane4 ; bytes 1..3 of lane4 are zero
    switch (k) {
        case 0: return __byte_perm(word, lane4, 0x5504);
        case 1: return __byte_perm(word, lane4, 0x5514);
        case 2: return __byte_perm(word, lane4, 0x5524);
        default: return __byte_perm(word, lane4, 0x5534);
    }
}
#define BV_REP_ADDR(W, k, lane4) rep_addr((W)[(k) >> 2], (lane4), (k)&3)

// ----------------------------------------------------------------------------------------------
// pass 2: table -> BGR2HSV -> S, V histograms; H,S,V kept in the L2-resident scratch image
// ----------------------------------------------------------------------------------------------
struct Sv2Smem {
    unsigned char rep[kRepBytes];          // word 0: packed BGR tables, word 1: sdiv
    int hdiv8[256 * 8];                    // entry i, copy c at word i*8 + c
    uint32_t h[kFastWarps][2][256];        // per-warp private S and V histograms
    uint8_t lut[3][256];                   // staging of the frame's tables
    int sdiv[256], hdiv[256];              // one copy, computed once per block, then replicated
    StatScratch sc;
};

// rep / hdiv8: the block's tables (uniform base addresses, folded into the load's immediate); lane4 = lane * 4,
// sd_off = lane * 4 + 128 (word 1 of this lane's slot), hd_off = (lane % 8) * 4
template <int J>
__device__ __forceinline__ void hsv_group_fast(const Px16 &in, const unsigned char *rep, const unsigned char *hdiv8,
                                               uint32_t (*hw)[256], uint32_t lane4, uint32_t sd_off, uint32_t hd_off, Px16 &o) {
    if constexpr (J < 16) {
        const int b = rep[BV_REP_ADDR(in.w, 3 * J, lane4) + 0];
        const int g = rep[BV_REP_ADDR(in.w, 3 * J + 1, lane4) + 1];
        const int r = rep[BV_REP_ADDR(in.w, 3 * J + 2, lane4) + 2];
        const int v = __vimax3_s32(b, g, r);
        const int diff = v - __vimin3_s32(b, g, r);
        const int hr = g - b, hg = b - r + 2 * diff, hb = r - g + 4 * diff;
        int hh = (v == g) ? hg : hb;
        hh = (v == r) ? hr : hh;
        const int sd = *reinterpret_cast<const int *>(rep + (((uint32_t)v << 8) + sd_off));
        const int hd = *reinterpret_cast<const int *>(hdiv8 + (((uint32_t)diff << 5) + hd_off));
        const int s = (diff * sd + (1 << (kHsvShift - 1))) >> kHsvShift;
        hh = (hh * hd + (1 << (kHsvShift - 1))) >> kHsvShift;
        const int h = hh + ((hh >> 31) & 180);
        atomicAdd(&hw[0][s], 1u);
        atomicAdd(&hw[1][v], 1u);
        put_px<J>(o.w, (uint32_t)h | ((uint32_t)s << 8) | ((uint32_t)v << 16));
        hsv_group_fast<J + 1>(in, rep, hdiv8, hw, lane4, sd_off, hd_off, o);
    }
}

__global__ void __launch_bounds__(kFastThreads, 2) hist_sv_fast_kernel(const uint8_t *__restrict__ src, BalFrame *__restrict__ st,
                                                                       size_t npx, uint8_t *__restrict__ hsv, size_t hsv_stride) {
    extern __shared__ __align__(128) unsigned char fast_raw[];
    Sv2Smem &sm = *reinterpret_cast<Sv2Smem *>(fast_raw);
    const int frame = blockIdx.y, t = threadIdx.x, lane = t & 31;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    // frame-independent part of the prologue (overlaps the previous pass under programmatic dependent launch)
    for (int i = t; i < kFastWarps * 512; i += kFastThreads) (&sm.h[0][0][0])[i] = 0;
    if (t < 256) {
        sm.sdiv[t] = hsv_sdiv(t);
        sm.hdiv[t] = hsv_hdiv(t);
    }
    __syncthreads();
    for (int i = t; i < 256 * 8; i += kFastThreads) sm.hdiv8[i] = sm.hdiv[i >> 3];
    for (int i = t; i < 256 * 32; i += kFastThreads)  // word 1 of row i/32, lane i%32
        *reinterpret_cast<int *>(sm.rep + ((i >> 5) << 8) + 128 + ((i & 31) << 2)) = sm.sdiv[i >> 5];
    grid_dependency_wait();  // pass 1 (the tables of this frame) is complete from here on
    for (int i = t; i < 768; i += kFastThreads) (&sm.lut[0][0])[i] = (&st[frame].lut_bgr[0][0])[i];
    __syncthreads();
    for (int i = t; i < 256 * 32; i += kFastThreads) {
        const int x = i >> 5;
        *reinterpret_cast<uint32_t *>(sm.rep + (x << 8) + ((i & 31) << 2)) =
            (uint32_t)sm.lut[0][x] | ((uint32_t)sm.lut[1][x] << 8) | ((uint32_t)sm.lut[2][x] << 16);
    }
    __syncthreads();
    const uint8_t *f = src + (size_t)frame * npx * 3;
    uint32_t(*hw)[256] = sm.h[t >> 5];
    const unsigned char *hdiv8 = reinterpret_cast<const unsigned char *>(sm.hdiv8);
    const uint32_t sd_off = lane4 + 128u, hd_off = (uint32_t)(lane & 7) << 2;
    const size_t stride = (size_t)gridDim.x * kFastThreads;
    const size_t ngroups = npx / 16;
    uint8_t *hf = hsv + (size_t)frame * hsv_stride;
    size_t g = (size_t)blockIdx.x * kFastThreads + t;
    Px16 in, nxt;
    if (g < ngroups) load_px16<true>(f, g, in);
    for (; g < ngroups; g += stride) {
        if (g + stride < ngroups) load_px16<true>(f, g + stride, nxt);
        Px16 o;
#pragma unroll
        for (int k = 0; k < 12; ++k) o.w[k] = 0;
        hsv_group_fast<0>(in, sm.rep, hdiv8, hw, lane4, sd_off, hd_off, o);
        store_px16_keep(hf, g, o);
        in = nxt;
    }
    __syncthreads();
    for (int i = t; i < 512; i += kFastThreads) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kFastWarps; ++w) s += (&sm.h[w][0][0])[i];
        if (s) atomicAdd(&st[frame].hist_sv[0][0] + i, s);
    }
    if (last_block_of_frame(&st[frame].ticket[1], gridDim.x) && t < 256) stats_sv_block(st[frame], npx, sm.sc);
}

// ----------------------------------------------------------------------------------------------
// pass 3 from the H,S,V scratch: S/V stretch -> HSV2BGR -> [convert -> inRange] -> outputs
// ----------------------------------------------------------------------------------------------
struct Fin2Smem {
    unsigned char rep[kRepBytes];   // word 0: float(2^23 + S'[x]), word 1: float(2^23 + V'[x])
    int gamma32[256 * 32];          // Lab: gamma table replicated per lane, entry i of lane l at word i*32 + l
    uint8_t lsv[2][256];
    SmemTabs tabs;                  // cube-root table (Lab), sdiv / hdiv (HSV)
};
// without the 32 KB gamma replica (every conversion but Lab): same leading layout
struct Fin2SmemSmall {
    unsigned char rep[kRepBytes];
    uint8_t lsv[2][256];
    SmemTabs tabs;
};

// HSV -> BGR of pixel_math.cuh with the three inputs already in biased-float form: hb = bits of float(2^23 + H),
// sb / vb = bits of float(2^23 + S') / float(2^23 + V').  Vector-path rounding only (whole 32-pixel groups of a row).
__device__ __forceinline__ uint32_t hsv2bgr_biased(uint32_t hb, uint32_t sb, uint32_t vb) {
    const float hscale = 6.f / 180.f;
    const float inv255 = 1.f / 255.f;
    const float kTwo23 = 8388608.f;
    const float h = __fmaf_rn(__uint_as_float(hb), hscale, -kTwo23 * hscale);
    const float s = __fmaf_rn(__uint_as_float(sb), inv255, -kTwo23 * inv255);
    const float v = __fmaf_rn(__uint_as_float(vb), inv255, -kTwo23 * inv255);
    const float hfloor = __fadd_rz(h, kTwo23);
    const int sector = (int)(__float_as_uint(hfloor) & 15u);
    const float f = __fsub_rn(h, __fsub_rn(hfloor, kTwo23));
    const float fm = (sector & 1) ? f : __fsub_rn(1.f, f);
    const float ymax = __fmul_rn(v, 255.f);
    const float ymin = __fmul_rn(__fmul_rn(v, __fsub_rn(1.f, s)), 255.f);
    const float ymid = __fmul_rn(__fmul_rn(v, __fmaf_rn(-s, fm, 1.f)), 255.f);
    const uint32_t amax = __float_as_uint(__fadd_rz(ymax, kTwo23));
    const uint32_t amid = __float_as_uint(__fadd_rz(ymid, kTwo23));
    const uint32_t amin = __float_as_uint(__fadd_rz(ymin, kTwo23));
    const uint32_t w = __byte_perm(__byte_perm(amax, amid, 0x2240), amin, 0x3410);
    const unsigned long long kSel = 0x012ull | (0x102ull << 10) | (0x201ull << 20) | (0x210ull << 30) | (0x120ull << 40) |
                                    (0x021ull << 50);
    const uint32_t sel = (uint32_t)(kSel >> (10 * sector)) & 0x3FFu;
    return __byte_perm(w, 0u, sel | 0x4000u);
}

// BGR -> Lab with the gamma look-ups from the per-lane replica (entry i of lane l at byte i*128 + l*4); p = b | g << 8 | r << 16
__device__ __forceinline__ uint32_t bgr2lab_fast(uint32_t p, const unsigned char *gamma, uint32_t lane4, const uint16_t *ctab) {
    const int B = *reinterpret_cast<const int *>(gamma + (((p << 7) & 0x7F80u) | lane4));
    const int G = *reinterpret_cast<const int *>(gamma + (((p >> 1) & 0x7F80u) | lane4));
    const int R = *reinterpret_cast<const int *>(gamma + (((p >> 9) & 0x7F80u) | lane4));
    const int fX = ctab[descale(R * 1777 + G * 1541 + B * 778, 12)];
    const int fY = ctab[descale(R * 871 + G * 2929 + B * 296, 12)];
    const int fZ = ctab[descale(R * 73 + G * 448 + B * 3575, 12)];
    const int L = descale(296 * fY - 1336934, 15);
    const int a = descale(500 * (fX - fY) + (128 << 15), 15);
    const int bb = descale(200 * (fY - fZ) + (128 << 15), 15);
    return (uint32_t)L | ((uint32_t)a << 8) | ((uint32_t)bb << 16);
}

template <int CODE, bool NEED_MASK, int J>
struct FastGroup {
    static __device__ __forceinline__ void run(const Px16 &in, const unsigned char *rep, const unsigned char *gamma, uint32_t lane4,
                                               const SmemTabs &tabs, const RangeTest &bd, Px16 &ob, Px16 &oc, uint32_t (&q)[4],
                                               uint32_t &bits) {
        constexpr bool kOne = CvtTraits<CODE>::kOneChannel;
        const uint32_t hb = (J * 3) % 4 == 0   ? __byte_perm(in.w[(3 * J) >> 2], 0x4B000000u, 0x7650)
                            : (J * 3) % 4 == 1 ? __byte_perm(in.w[(3 * J) >> 2], 0x4B000000u, 0x7651)
                            : (J * 3) % 4 == 2 ? __byte_perm(in.w[(3 * J) >> 2], 0x4B000000u, 0x7652)
                                               : __byte_perm(in.w[(3 * J) >> 2], 0x4B000000u, 0x7653);
        const uint32_t sb = *reinterpret_cast<const uint32_t *>(rep + BV_REP_ADDR(in.w, 3 * J + 1, lane4));
        const uint32_t vb = *reinterpret_cast<const uint32_t *>(rep + BV_REP_ADDR(in.w, 3 * J + 2, lane4) + 128);
        const uint32_t p = hsv2bgr_biased(hb, sb, vb);
        put_px<J>(ob.w, p);
        uint32_t pc;
        int o0, o1, o2;
        if (CODE == BV_BGR2LAB) {
            pc = bgr2lab_fast(p, gamma, lane4, tabs.ctab);
            o0 = (int)(pc & 0xFF);
            o1 = (int)((pc >> 8) & 0xFF);
            o2 = (int)(pc >> 16);
        } else {
            convert_px<CODE>((int)(p & 0xFF), (int)((p >> 8) & 0xFF), (int)(p >> 16), true, tabs, o0, o1, o2);
            pc = (uint32_t)o0 | ((uint32_t)o1 << 8) | ((uint32_t)o2 << 16);
        }
        if (kOne)
            BV_PUTB(q, J, o0);
        else
            put_px<J>(oc.w, pc);
        if (NEED_MASK)
            if (in_range_px<CODE>(o0, o1, o2, bd)) bits |= 1u << J;
        FastGroup<CODE, NEED_MASK, J + 1>::run(in, rep, gamma, lane4, tabs, bd, ob, oc, q, bits);
    }
};
template <int CODE, bool NEED_MASK>
struct FastGroup<CODE, NEED_MASK, 16> {
    static __device__ __forceinline__ void run(const Px16 &, const unsigned char *, const unsigned char *, uint32_t, const SmemTabs &,
                                               const RangeTest &, Px16 &, Px16 &, uint32_t (&)[4], uint32_t &) {}
};

// Requires: width % 32 == 0 (vector-path rounding everywhere), npx % 16 == 0, 16-byte aligned buffers.
template <int CODE, bool NEED_MASK>
__global__ void __launch_bounds__(kFastThreads, 2) final_fast_kernel(const uint8_t *__restrict__ hsv, size_t hsv_stride,
                                                                     const BalFrame *__restrict__ st, size_t npx, int width,
                                                                     BalOutputs out, const uint16_t *__restrict__ g_gamma,
                                                                     const uint16_t *__restrict__ g_cbrt) {
    extern __shared__ __align__(128) unsigned char fast_raw[];
    constexpr bool kLab = CODE == BV_BGR2LAB;
    // the two layouts share the leading `rep`; everything else is addressed through these pointers
    unsigned char *rep = fast_raw;
    int *gamma32 = kLab ? reinterpret_cast<Fin2Smem *>(fast_raw)->gamma32 : nullptr;
    uint8_t(*lsv)[256] = kLab ? reinterpret_cast<Fin2Smem *>(fast_raw)->lsv : reinterpret_cast<Fin2SmemSmall *>(fast_raw)->lsv;
    SmemTabs &tabs = kLab ? reinterpret_cast<Fin2Smem *>(fast_raw)->tabs : reinterpret_cast<Fin2SmemSmall *>(fast_raw)->tabs;
    const int frame = blockIdx.y, t = threadIdx.x, lane = t & 31;
    const uint32_t lane4 = (uint32_t)lane * 4u;
    if (kLab) {
        for (int i = t; i < kLabCbrtSize; i += kFastThreads) tabs.ctab[i] = g_cbrt[i];
        for (int i = t; i < 256 * 32; i += kFastThreads) gamma32[i] = g_gamma[i >> 5];
    } else {
        init_tabs<CODE>(tabs, g_gamma, g_cbrt);
    }
    grid_dependency_wait();  // pass 2 (S/V tables of this frame, the H,S,V scratch) is complete from here on
    for (int i = t; i < 512; i += kFastThreads) (&lsv[0][0])[i] = (&st[frame].lut_sv[0][0])[i];
    __syncthreads();
    for (int i = t; i < 256 * 64; i += kFastThreads) {  // word i%64 of row i/64: lanes 0..31 word 0, then word 1
        const int x = i >> 6, c = (i >> 5) & 1;
        *reinterpret_cast<uint32_t *>(rep + (x << 8) + ((i & 63) << 2)) = 0x4B000000u | lsv[c][x];
    }
    __syncthreads();
    const size_t foff = (size_t)frame * npx;
    const uint8_t *f = hsv + (size_t)frame * hsv_stride;
    constexpr bool kOne = CvtTraits<CODE>::kOneChannel;
    const RangeTest bd = make_range_test(out.lo, out.hi);
    const uint32_t stride = gridDim.x * kFastThreads;
    const uint32_t ngroups = (uint32_t)(npx / 16);
    const uint32_t height = (uint32_t)(npx / (size_t)width);
    const int wp2 = ((width + 31) / 32) * 2;
    const unsigned char *gamma = reinterpret_cast<const unsigned char *>(gamma32);
    for (uint32_t g = blockIdx.x * kFastThreads + t; g < ngroups; g += stride) {
        Px16 in;
        load_px16<false>(f, g, in);
        Px16 ob, oc;
        uint32_t q[4] = {0, 0, 0, 0};
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 12; ++k) ob.w[k] = oc.w[k] = 0;
        FastGroup<CODE, NEED_MASK, 0>::run(in, rep, gamma, lane4, tabs, bd, ob, oc, q, bits);
        if (out.balanced) store_px16(out.balanced + foff * 3, g, ob);
        if (out.converted) {
            if (kOne)
                st_stream(reinterpret_cast<uint4 *>(out.converted + foff) + g, make_uint4(q[0], q[1], q[2], q[3]));
            else
                store_px16(out.converted + foff * 3, g, oc);
        }
        if (NEED_MASK && out.mask) {
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t nib = (bits >> (4 * k)) & 0xF;
                m[k] = ((nib & 1) * 0xFFu) | (((nib >> 1) & 1) * 0xFF00u) | (((nib >> 2) & 1) * 0xFF0000u) |
                       (((nib >> 3) & 1) * 0xFF000000u);
            }
            st_stream(reinterpret_cast<uint4 *>(out.mask + foff) + g, make_uint4(m[0], m[1], m[2], m[3]));
        }
        if (NEED_MASK && out.mask_bits) {
            const uint32_t p0 = g * 16u, y = p0 / (uint32_t)width, x0 = p0 - y * (uint32_t)width;
            out.mask_bits[((size_t)frame * height + y) * wp2 + (x0 >> 4)] = (uint16_t)bits;
        }
    }
}

}  // namespace bv
