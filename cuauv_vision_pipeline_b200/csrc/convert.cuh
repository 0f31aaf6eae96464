// convert.cuh -- shared-memory tables and the per-pixel conversion switch used by every kernel
// that converts colour spaces (cvt.cu, balance.cu).
#pragma once
#include "common.cuh"

namespace bv {

struct SmemTabs {
    int sdiv[256];
    int hdiv[256];
    uint16_t gtab[kLabGammaSize];
    uint16_t ctab[kLabCbrtSize];  // BV_LAB2BGR reuses these 6 KB for its own tables (lab_yf / lab_inv_gamma below)
};
// Lab -> BGR tables live in the ctab area: 256 x 2 u16 (y, fy) followed by the 4096-byte inverse gamma
__device__ __forceinline__ const uint16_t *lab_yf(const SmemTabs &t) { return t.ctab; }
__device__ __forceinline__ const uint8_t *lab_inv_gamma(const SmemTabs &t) { return reinterpret_cast<const uint8_t *>(t.ctab + 512); }
static_assert(512 * 2 + kLabInvGammaSize <= kLabCbrtSize * 2, "Lab->BGR tables must fit the cube-root table's shared memory");

template <int CODE>
__device__ __forceinline__ void init_tabs(SmemTabs &t, const uint16_t *__restrict__ g_gamma,
                                          const uint16_t *__restrict__ g_cbrt) {
    if (CODE == BV_BGR2LAB) {
        for (int i = threadIdx.x; i < kLabGammaSize; i += blockDim.x) t.gtab[i] = g_gamma[i];
        for (int i = threadIdx.x; i < kLabCbrtSize; i += blockDim.x) t.ctab[i] = g_cbrt[i];
    }
    if (CODE == BV_LAB2BGR)  // the device buffer behind g_cbrt continues with (y, fy) and the inverse gamma table
        for (int i = threadIdx.x; i < 512 + kLabInvGammaSize / 2; i += blockDim.x) t.ctab[i] = g_cbrt[kLabCbrtSize + i];
    __syncthreads();
}

// One pixel through conversion CODE.  `vec` is the cv2 vector-path flag (HSV2BGR / HLS only).
template <int CODE>
__device__ __forceinline__ void convert_px(int c0, int c1, int c2, bool vec, const SmemTabs &t, int &o0, int &o1,
                                           int &o2) {
    if (CODE == BV_BGR2HSV) {
#if defined(__CUDA_ARCH__)
        bgr2hsv_rcp(c0, c1, c2, o0, o1, o2);   // sdiv / hdiv from the reciprocal unit: identical values (checked per device at bv_create)
#else
        bgr2hsv(c0, c1, c2, t.sdiv, t.hdiv, o0, o1, o2);
#endif
    } else if (CODE == BV_BGR2LAB) {
        bgr2lab(c0, c1, c2, t.gtab, t.ctab, o0, o1, o2);
    } else if (CODE == BV_BGR2GRAY) {
        o0 = bgr2gray(c0, c1, c2);
        o1 = o2 = 0;
    } else if (CODE == BV_BGR2YCRCB) {
        bgr2ycrcb(c0, c1, c2, o0, o1, o2);
    } else if (CODE == BV_HSV2BGR) {
        hsv2bgr(c0, c1, c2, vec, o0, o1, o2);
    } else if (CODE == BV_BGR2HLS) {
        bgr2hls(c0, c1, c2, vec, o0, o1, o2);
    } else if (CODE == BV_LAB2BGR) {
        lab2bgr(c0, c1, c2, lab_yf(t), lab_inv_gamma(t), o0, o1, o2);
    } else if (CODE == BV_BGR2RGB) {
        o0 = c2;
        o1 = c1;
        o2 = c0;
    } else {  // identity (-1)
        o0 = c0;
        o1 = c1;
        o2 = c2;
    }
}

template <int CODE>
struct CvtTraits {
    static constexpr bool kOneChannel = (CODE == BV_BGR2GRAY);
    static constexpr bool kNeedsX = (CODE == BV_HSV2BGR || CODE == BV_BGR2HLS);
};

struct Bounds3 {
    uint8_t lo[3], hi[3];
};

// inclusive range test as one subtract + one unsigned compare per channel:
//   lo <= x <= hi  <=>  (unsigned)(x - lo) <= (unsigned)(hi - lo);   an empty range (hi < lo) is
// encoded as lo = 256, span = 0 so that no 8-bit value passes.
struct RangeTest {
    int lo[3];
    unsigned span[3];
};

__host__ __device__ inline RangeTest make_range_test(const uint8_t *lo, const uint8_t *hi) {
    RangeTest r;
    for (int k = 0; k < 3; ++k) {
        if (hi[k] < lo[k]) {
            r.lo[k] = 256;
            r.span[k] = 0;
        } else {
            r.lo[k] = lo[k];
            r.span[k] = (unsigned)(hi[k] - lo[k]);
        }
    }
    return r;
}

template <int CODE>
__device__ __forceinline__ bool in_range_px(int o0, int o1, int o2, const RangeTest &rt) {
    bool in_r = (unsigned)(o0 - rt.lo[0]) <= rt.span[0];
    if (!CvtTraits<CODE>::kOneChannel)
        in_r = in_r && (unsigned)(o1 - rt.lo[1]) <= rt.span[1] && (unsigned)(o2 - rt.lo[2]) <= rt.span[2];
    return in_r;
}

template <int CODE>
__device__ __forceinline__ bool in_range_px(int o0, int o1, int o2, const Bounds3 &bd) {
    bool in_r = o0 >= bd.lo[0] && o0 <= bd.hi[0];
    if (!CvtTraits<CODE>::kOneChannel) in_r = in_r && o1 >= bd.lo[1] && o1 <= bd.hi[1] && o2 >= bd.lo[2] && o2 <= bd.hi[2];
    return in_r;
}

}  // namespace bv
