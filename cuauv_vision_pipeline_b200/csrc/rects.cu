// rects.cu -- cv2.minAreaRect of every outer contour (modules/bins.py:60-69, utils/feature.py:301-312),
// on the device-side CHAIN_APPROX_SIMPLE vertex lists that bv_outer_contours leaves in HBM, so the
// module's rectangle filter needs no per-contour host round trip.
//
// One thread per contour (the point sets are tiny: tens to hundreds of vertices):
//   1. heap-sort a private copy of the vertices by (x, y); drop duplicates
//   2. Andrew's monotone chain with exact integer cross products -> convex hull, taken clockwise in
//      y-up axes (the orientation cv::minAreaRect feeds to its rotating calipers)
//   3. rotating calipers in float32 as OpenCV's rotatingCalipers(CALIPERS_MINAREARECT) does: the
//      base vector follows the hull edge with the smallest angle to one of the four caliper sides;
//      the rectangle of least area (ties: the last one met) is kept
//   4. centre / size / angle as cv::minAreaRect derives them, angle reported in [-90, 0) like
//      cv2 4.13.0 (quarter turns with the sides swapped until it is in range)
// Tolerance (stated, tests/test_gpu_morph_ccl.py): the area agrees with cv2 to 1e-4 relative; centre,
// size and angle to 1e-3 whenever the minimum is unique -- OpenCV's own float32 evaluation order is
// not reproduced bit for bit, so two hull edges whose rectangles tie to within rounding may swap.
#include <math.h>

#include "common.cuh"

namespace bv {

struct P2 {
    int x, y;
};

__device__ __forceinline__ bool p_less(const P2 &a, const P2 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

__device__ void heap_sort(P2 *a, int n) {
    for (int start = n / 2 - 1; start >= 0; --start) {  // heapify
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && p_less(a[child], a[child + 1])) ++child;
            if (!p_less(a[root], a[child])) break;
            const P2 t = a[root]; a[root] = a[child]; a[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        const P2 t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && p_less(a[child], a[child + 1])) ++child;
            if (!p_less(a[root], a[child])) break;
            const P2 u = a[root]; a[root] = a[child]; a[child] = u;
            root = child;
        }
    }
}

__device__ __forceinline__ long long cross3(const P2 &o, const P2 &a, const P2 &b) {
    return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// hull vertex i of the clockwise (y-up) sequence = counter-clockwise sequence reversed
struct HullView {
    const P2 *h;
    int n;
    const float *ex, *ey, *el;  // edge vectors and inverse lengths computed once per edge (NULL: computed on demand)
    __device__ __forceinline__ float2 pt(int i) const {
        const P2 p = h[n - 1 - i];
        return make_float2((float)p.x, (float)p.y);
    }
    __device__ __forceinline__ void edge(int i, float &vx, float &vy, float &inv_len) const {
        if (ex) {
            vx = ex[i];
            vy = ey[i];
            inv_len = el[i];
            return;
        }
        compute_edge(i, vx, vy, inv_len);
    }
    __device__ __forceinline__ void compute_edge(int i, float &vx, float &vy, float &inv_len) const {
        const P2 a = h[n - 1 - i], b = h[n - 1 - (i + 1 == n ? 0 : i + 1)];
        const double dx = (double)b.x - (double)a.x, dy = (double)b.y - (double)a.y;
        vx = (float)dx;
        vy = (float)dy;
        inv_len = (float)(1. / sqrt(dx * dx + dy * dy));
    }
};

__device__ void rotating_calipers(const HullView &H, float out[6]) {
    const int n = H.n;
    int left = 0, bottom = 0, right = 0, top = 0;
    float2 p0 = H.pt(0);
    float left_x = p0.x, right_x = p0.x, top_y = p0.y, bottom_y = p0.y;
    for (int i = 0; i < n; ++i) {
        const float2 p = H.pt(i);
        if (p.x < left_x) { left_x = p.x; left = i; }
        if (p.x > right_x) { right_x = p.x; right = i; }
        if (p.y > top_y) { top_y = p.y; top = i; }
        if (p.y < bottom_y) { bottom_y = p.y; bottom = i; }
    }
    // orientation of the hull: sign of the first non-zero turn
    float orientation = 0.f;
    {
        float ax, ay, il;
        H.edge(n - 1, ax, ay, il);
        for (int i = 0; i < n; ++i) {
            float bx, by;
            H.edge(i, bx, by, il);
            const double convexity = (double)ax * by - (double)ay * bx;
            if (convexity != 0) {
                orientation = convexity > 0 ? 1.f : -1.f;
                break;
            }
            ax = bx;
            ay = by;
        }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    float minarea = 3.402823466e+38f;
    int best_left = 0, best_bottom = 0;
    float best_a = 1.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
    for (int k = 0; k < n; ++k) {
        float vx[4], vy[4], il[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) H.edge(seq[q], vx[q], vy[q], il[q]);
        const float dp[4] = {+base_a * vx[0] + base_b * vy[0], -base_b * vx[1] + base_a * vy[1],
                             -base_a * vx[2] - base_b * vy[2], +base_b * vx[3] - base_a * vy[3]};
        float maxcos = dp[0] * il[0];
        int main_el = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const float c = dp[q] * il[q];
            if (c > maxcos) {
                main_el = q;
                maxcos = c;
            }
        }
        float lx = 0.f, ly = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q == main_el) {
                lx = vx[q] * il[q];
                ly = vy[q] * il[q];
            }
        if (main_el == 0) { base_a = lx; base_b = ly; }
        else if (main_el == 1) { base_a = ly; base_b = -lx; }
        else if (main_el == 2) { base_a = -lx; base_b = -ly; }
        else { base_a = -ly; base_b = lx; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q == main_el) seq[q] = seq[q] + 1 == n ? 0 : seq[q] + 1;
        const float2 pr = H.pt(seq[1]), pl = H.pt(seq[3]), pt_ = H.pt(seq[2]), pb = H.pt(seq[0]);
        float dx = pr.x - pl.x, dy = pr.y - pl.y;
        const float width = dx * base_a + dy * base_b;
        dx = pt_.x - pb.x;
        dy = pt_.y - pb.y;
        const float height = -dx * base_b + dy * base_a;
        const float area = width * height;
        if (area <= minarea) {
            minarea = area;
            best_left = seq[3];
            best_bottom = seq[0];
            best_a = base_a;
            best_b = base_b;
            best_w = width;
            best_h = height;
        }
    }
    const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
    const float2 pl = H.pt(best_left), pb = H.pt(best_bottom);
    const float C1 = A1 * pl.x + pl.y * B1;
    const float C2 = A2 * pb.x + pb.y * B2;
    const float idet = 1.f / (A1 * B2 - A2 * B1);
    out[0] = (C1 * B2 - C2 * B1) * idet;
    out[1] = (A1 * C2 - A2 * C1) * idet;
    out[2] = A1 * best_w;
    out[3] = B1 * best_w;
    out[4] = A2 * best_h;
    out[5] = B2 * best_h;
}

// Akl-Toussaint prefilter + heap sort + monotone chain by ONE thread in global scratch: the fallback for contours whose
// kept vertices do not fit the warp's shared-memory buffer (or whose coordinates do not fit 16 bits).  Returns the hull size.
__device__ __noinline__ int hull_sequential(const P2 *__restrict__ src, int n_in, P2 *sorted, P2 *hull) {
    P2 ext[4] = {src[0], src[0], src[0], src[0]};
    for (int i = 1; i < n_in; ++i) {
        const P2 p = src[i];
        if (p.x < ext[0].x) ext[0] = p;
        if (p.y < ext[1].y) ext[1] = p;
        if (p.x > ext[2].x) ext[2] = p;
        if (p.y > ext[3].y) ext[3] = p;
    }
    long long area2 = 0;
    for (int k = 0; k < 4; ++k)
        area2 += (long long)ext[k].x * ext[(k + 1) & 3].y - (long long)ext[(k + 1) & 3].x * ext[k].y;
    int n_kept = 0;
    if (area2 == 0 || n_in <= 8) {
        for (int i = 0; i < n_in; ++i) sorted[n_kept++] = src[i];
    } else {
        for (int k = 0; k < 4; ++k) sorted[n_kept++] = ext[k];
        for (int i = 0; i < n_in; ++i) {
            const P2 p = src[i];
            bool outside = false;
            for (int k = 0; k < 4; ++k) {
                const long long c = cross3(ext[k], ext[(k + 1) & 3], p);
                outside |= area2 > 0 ? c < 0 : c > 0;
            }
            if (outside) sorted[n_kept++] = p;
        }
    }
    heap_sort(sorted, n_kept);
    int n = 0;
    for (int i = 0; i < n_kept; ++i)
        if (n == 0 || sorted[i].x != sorted[n - 1].x || sorted[i].y != sorted[n - 1].y) sorted[n++] = sorted[i];
    int m = 0;
    if (n <= 2) {
        for (int i = 0; i < n; ++i) hull[m++] = sorted[i];
    } else {
        for (int i = 0; i < n; ++i) {  // lower chain
            while (m >= 2 && cross3(hull[m - 2], hull[m - 1], sorted[i]) <= 0) --m;
            hull[m++] = sorted[i];
        }
        const int lower = m + 1;
        for (int i = n - 2; i >= 0; --i) {  // upper chain
            while (m >= lower && cross3(hull[m - 2], hull[m - 1], sorted[i]) <= 0) --m;
            hull[m++] = sorted[i];
        }
        --m;  // the first point again
    }
    return m;
}

// cv2.minAreaRect of a convex hull (counter-clockwise in hull[0..m), as the monotone chain leaves it)
__device__ void rect_from_hull(const HullView &H, bv_rrect &r) {
    const P2 *hull = H.h;
    const int m = H.n;
    float angle = 0.f;
    if (m > 2) {
        float o[6];
        rotating_calipers(H, o);
        r.cx = o[0] + (o[2] + o[4]) * 0.5f;
        r.cy = o[1] + (o[3] + o[5]) * 0.5f;
        r.width = (float)sqrt((double)o[2] * o[2] + (double)o[3] * o[3]);
        r.height = (float)sqrt((double)o[4] * o[4] + (double)o[5] * o[5]);
        angle = (float)atan2((double)o[3], (double)o[2]);
    } else if (m == 2) {
        const P2 a = hull[0], b = hull[1];  // (x, y)-sorted, the order cv::convexHull returns a segment in
        r.cx = ((float)a.x + (float)b.x) * 0.5f;
        r.cy = ((float)a.y + (float)b.y) * 0.5f;
        const double dx = (double)b.x - a.x, dy = (double)b.y - a.y;
        r.width = (float)sqrt(dx * dx + dy * dy);
        r.height = 0.f;
        angle = (float)atan2(dy, dx);
    } else {
        r.cx = (float)hull[0].x;
        r.cy = (float)hull[0].y;
    }
    angle = (float)((double)angle * 180. / 3.1415926535897932384626433832795);
    // cv2 4.13.0 reports the angle in [-90, 0): quarter turns, the sides swapping with each
    for (int q = 0; q < 4 && !(angle < 0.f); ++q) {
        angle -= 90.f;
        const float t = r.width; r.width = r.height; r.height = t;
    }
    for (int q = 0; q < 4 && angle < -90.f; ++q) {
        angle += 90.f;
        const float t = r.width; r.width = r.height; r.height = t;
    }
    r.angle = angle;
    r.valid = 1;
}

// One WARP per contour.  A thread per contour spends its time in a heap sort over global scratch (one dependent access
// after the other, the frame waits for its largest contour); here the 32 lanes find the extreme points, filter and
// compact the vertices into shared memory (ballot), sort them as packed 16+16-bit keys with a bitonic network, and
// compute the hull's edge vectors; only the monotone chain and the calipers, both O(hull), run on lane 0, out of
// shared memory.
constexpr int kRectCap = 512;      // vertices a warp keeps in shared memory; more (or coordinates > 65534) -> fallback
constexpr int kRectWarps = 2;      // warps per block

struct RectWarpSmem {
    uint32_t keys[kRectCap];
    P2 hull[kRectCap + 1];
    float ex[kRectCap], ey[kRectCap], el[kRectCap];
};

__device__ __forceinline__ P2 key_point(uint32_t k) { return P2{(int)(k >> 16), (int)(k & 0xFFFFu)}; }

__global__ void __launch_bounds__(32 * kRectWarps) min_area_rect_kernel(const bv_contour *__restrict__ contours,
                                                                        const int32_t *__restrict__ n_contours,
                                                                        const int32_t *__restrict__ points, int max_contours,
                                                                        int max_points, P2 *__restrict__ scratch,
                                                                        bv_rrect *__restrict__ rects) {
    __shared__ RectWarpSmem smem[kRectWarps];
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ci = blockIdx.x * kRectWarps + wib;
    if (ci >= max_contours) return;  // warp-uniform
    RectWarpSmem &sm = smem[wib];
    bv_rrect r;
    r.cx = r.cy = r.width = r.height = r.angle = 0.f;
    r.valid = 0;
    const int total = min(n_contours[frame], max_contours);
    const bv_contour c = contours[(size_t)frame * max_contours + ci];
    if (ci < total && c.external && c.point_offset >= 0 && c.n_simple > 0 && c.point_offset + c.n_simple <= max_points) {
        const int n_in = c.n_simple;
        const P2 *src = reinterpret_cast<const P2 *>(points) + (size_t)frame * max_points + c.point_offset;
        // extreme points: per lane, then across the warp (ties: any of the tied points will do)
        P2 ext[4] = {src[0], src[0], src[0], src[0]};
        for (int i = lane; i < n_in; i += 32) {
            const P2 p = src[i];
            if (p.x < ext[0].x) ext[0] = p;
            if (p.y < ext[1].y) ext[1] = p;
            if (p.x > ext[2].x) ext[2] = p;
            if (p.y > ext[3].y) ext[3] = p;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            P2 o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k].x = __shfl_down_sync(0xFFFFFFFFu, ext[k].x, off);
                o[k].y = __shfl_down_sync(0xFFFFFFFFu, ext[k].y, off);
            }
            if (o[0].x < ext[0].x) ext[0] = o[0];
            if (o[1].y < ext[1].y) ext[1] = o[1];
            if (o[2].x > ext[2].x) ext[2] = o[2];
            if (o[3].y > ext[3].y) ext[3] = o[3];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ext[k].x = __shfl_sync(0xFFFFFFFFu, ext[k].x, 0);
            ext[k].y = __shfl_sync(0xFFFFFFFFu, ext[k].y, 0);
        }
        long long area2 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            area2 += (long long)ext[k].x * ext[(k + 1) & 3].y - (long long)ext[(k + 1) & 3].x * ext[k].y;
        // Akl-Toussaint: only the extremes and the vertices strictly outside their quadrilateral can be hull vertices
        const bool keep_all = area2 == 0 || n_in <= 8;
        int n_kept = 0;
        if (!keep_all) {
            if (lane < 4) sm.keys[lane] = ((uint32_t)ext[lane].x << 16) | (uint32_t)ext[lane].y;
            n_kept = 4;
        }
        bool wide = false;
        for (int base = 0; base < n_in; base += 32) {
            const int i = base + lane;
            bool keep = false;
            P2 p = P2{0, 0};
            if (i < n_in) {
                p = src[i];
                wide |= (unsigned)p.x > 65534u || (unsigned)p.y > 65534u;
                keep = keep_all;
                if (!keep_all) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const long long cr = cross3(ext[k], ext[(k + 1) & 3], p);
                        keep |= area2 > 0 ? cr < 0 : cr > 0;
                    }
                }
            }
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, keep);
            const int pos = n_kept + __popc(mask & ((1u << lane) - 1u));
            if (keep && pos < kRectCap) sm.keys[pos] = ((uint32_t)p.x << 16) | (uint32_t)p.y;
            n_kept += __popc(mask);
        }
        wide = __any_sync(0xFFFFFFFFu, wide);
        int m = 0;
        const P2 *hull = sm.hull;
        const bool fallback = wide || n_kept > kRectCap;  // warp-uniform
        if (fallback) {
            if (lane == 0) {
                // private scratch: [sorted points (n_in + 4) | hull stack (one more: the chain closes on its first point)]
                P2 *sorted = scratch + (size_t)frame * (2 * (size_t)max_points + 9 * (size_t)max_contours) + 2 * (size_t)c.point_offset +
                             9 * (size_t)ci;
                m = hull_sequential(src, n_in, sorted, sorted + n_in + 4);
                hull = sorted + n_in + 4;
            }
        } else {
            int np2 = 32;
            while (np2 < n_kept) np2 <<= 1;
            for (int i = n_kept + lane; i < np2; i += 32) sm.keys[i] = 0xFFFFFFFFu;  // above every real key
            __syncwarp();
            for (int k = 2; k <= np2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < (np2 >> 1); t += 32) {
                        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                        const uint32_t a = sm.keys[lo], b = sm.keys[hi];
                        if ((a > b) == ((lo & k) == 0)) {
                            sm.keys[lo] = b;
                            sm.keys[hi] = a;
                        }
                    }
                    __syncwarp();
                }
            if (lane == 0) {
                int n = 0;
                for (int i = 0; i < n_kept; ++i)
                    if (n == 0 || sm.keys[i] != sm.keys[n - 1]) sm.keys[n++] = sm.keys[i];
                if (n <= 2) {
                    for (int i = 0; i < n; ++i) sm.hull[m++] = key_point(sm.keys[i]);
                } else {
                    for (int i = 0; i < n; ++i) {  // lower chain
                        const P2 p = key_point(sm.keys[i]);
                        while (m >= 2 && cross3(sm.hull[m - 2], sm.hull[m - 1], p) <= 0) --m;
                        sm.hull[m++] = p;
                    }
                    const int lower = m + 1;
                    for (int i = n - 2; i >= 0; --i) {  // upper chain
                        const P2 p = key_point(sm.keys[i]);
                        while (m >= lower && cross3(sm.hull[m - 2], sm.hull[m - 1], p) <= 0) --m;
                        sm.hull[m++] = p;
                    }
                    --m;  // the first point again
                }
            }
        }
        m = __shfl_sync(0xFFFFFFFFu, m, 0);
        __syncwarp();
        HullView H{hull, m, nullptr, nullptr, nullptr};
        if (!fallback && m > 2) {  // edge vectors and inverse lengths once per edge, 32 at a time
            for (int i = lane; i < m; i += 32) H.compute_edge(i, sm.ex[i], sm.ey[i], sm.el[i]);
            H.ex = sm.ex;
            H.ey = sm.ey;
            H.el = sm.el;
        }
        __syncwarp();
        if (lane == 0) rect_from_hull(H, r);
    }
    if (lane == 0) rects[(size_t)frame * max_contours + ci] = r;
}

}  // namespace bv

using namespace bv;

extern "C" int bv_min_area_rects(bv_ctx *ctx, const bv_contour *contours_dev, const int32_t *n_contours_dev,
                                 const int32_t *points_dev, int batch, int max_contours, int max_points, bv_rrect *rects_dev) {
    BV_REQUIRE(ctx && contours_dev && n_contours_dev && points_dev && rects_dev, "null argument");
    BV_REQUIRE(batch > 0 && batch <= 65535 && max_contours > 0 && max_points > 0, "sizes must be positive");
    BV_CUDA(cudaSetDevice(ctx->device));
    BV_TRY(ensure_scratch(ctx, SCR_HULL, sizeof(P2) * (size_t)batch * (2 * (size_t)max_points + 9 * (size_t)max_contours)));
    BV_LAUNCH(ctx, min_area_rect_kernel, dim3((max_contours + kRectWarps - 1) / kRectWarps, batch), 32 * kRectWarps, 0, contours_dev,
              n_contours_dev, points_dev, max_contours, max_points, (P2 *)ctx->scratch[SCR_HULL], rects_dev);
    return BV_OK;
}
