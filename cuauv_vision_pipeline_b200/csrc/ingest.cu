// ingest.cu -- frames straight from a shared-memory ring guarded by a sequence lock to the device.
//
// The reference's transport keeps the last BUFFER_CNT frames of a camera in a POSIX shared-memory file
// (lib/camera_message_framework.cpp:39-54, struct Buffer): a writer copies the payload into slot (uid + 1) % slots,
// stamps the slot's v_a / v_b words and then publishes it by incrementing `uid` (:285-372).  Its own reader memcpy's the
// newest slot into a heap buffer and accepts the copy when v_a == v_b (:423-453); ModuleBase._loop then copies it once
// more into a writable array (core/base.py:762-768) before process() sees it, and a GPU module copies it a third time to
// the device.  bv_ingest_seqlock replaces all three by ONE host-to-device DMA out of the (registered) mapping, validated
// the same way after the copy has completed -- plus a lap check the reference does not make: the writer only stamps v_a
// AFTER it has overwritten the payload, so a reader whose slot is being rewritten can see v_a == v_b on torn bytes; the
// slot of the frame published as `uid0` is rewritten while uid == uid0 + slots - 1, hence any uid >= uid0 + slots - 1
// observed after the copy means "retry".  4-channel sources (the ZED hands out RGBA, capture_sources/zed.py:49-50,
// zed.cpp:54-71) have their alpha byte dropped by the first device kernel instead of a CPU loop on the capture side.
#include "common.cuh"

using namespace bv;

namespace bv {
int rgba_to_rgb_run(bv_ctx *ctx, const uint8_t *src_dev, uint8_t *dst_dev, size_t n_pixels, int swap_rb);  // aux.cu
}

static inline uint64_t load_acquire_u64(const volatile uint64_t *p) { return __atomic_load_n((const uint64_t *)p, __ATOMIC_ACQUIRE); }

extern "C" int bv_ingest_seqlock(bv_ctx *ctx, const bv_seqlock_ring *ring, size_t payload_offset, uint8_t *dst_dev, int height,
                                 int width, int src_channels, int swap_rb, int max_retries, uint64_t *uid_out, int *retries_out) {
    BV_REQUIRE(ctx && ring && dst_dev, "null argument");
    BV_REQUIRE(ring->uid && ring->v_begin && ring->v_end && ring->data && ring->slots >= 1, "incomplete ring description");
    BV_REQUIRE(height > 0 && width > 0, "height and width must be positive");
    BV_REQUIRE(src_channels == 3 || src_channels == 4, "src_channels must be 3 (BGR) or 4 (BGRA / RGBA)");
    BV_REQUIRE(!(swap_rb && src_channels == 3), "swap_rb needs a 4-channel source");
    const size_t npx = (size_t)height * width, bytes = npx * (size_t)src_channels;
    BV_REQUIRE(payload_offset + bytes <= ring->slot_stride, "frame does not fit a slot of the ring");
    BV_CUDA(cudaSetDevice(ctx->device));
    uint8_t *landing = dst_dev;
    if (src_channels == 4) {
        BV_TRY(ensure_scratch(ctx, SCR_INGEST, bytes));
        landing = (uint8_t *)ctx->scratch[SCR_INGEST];
    }
    int tries = 0;
    for (;; ++tries) {
        const uint64_t uid0 = load_acquire_u64(ring->uid);
        if (uid0 == 0) {
            set_error("bv_ingest_seqlock: nothing has been published yet");
            return BV_ERR_NOT_READY;
        }
        const size_t slot = (size_t)(uid0 % (uint64_t)ring->slots);
        const volatile uint64_t *vb = (const volatile uint64_t *)((const char *)ring->v_end + slot * ring->meta_stride);
        const volatile uint64_t *va = (const volatile uint64_t *)((const char *)ring->v_begin + slot * ring->meta_stride);
        const uint64_t v_end = load_acquire_u64(vb);
        BV_CUDA(cudaMemcpyAsync(landing, ring->data + slot * ring->slot_stride + payload_offset, bytes, cudaMemcpyHostToDevice,
                                ctx->stream));
        BV_CUDA(cudaStreamSynchronize(ctx->stream));  // the DMA has read every byte before the slot is re-validated
        const uint64_t v_begin = load_acquire_u64(va);
        const uint64_t uid1 = load_acquire_u64(ring->uid);
        const bool lapped = ring->slots > 1 ? uid1 + 1 >= uid0 + (uint64_t)ring->slots : uid1 != uid0;
        if (v_begin == v_end && !lapped) {
            if (uid_out) *uid_out = uid0;
            break;
        }
        if (tries >= max_retries) {
            if (retries_out) *retries_out = tries + 1;
            set_error("bv_ingest_seqlock: the writer lapped the reader %d times in a row", tries + 1);
            return BV_ERR_NOT_READY;
        }
    }
    if (retries_out) *retries_out = tries;
    if (src_channels == 4) BV_TRY(rgba_to_rgb_run(ctx, landing, dst_dev, npx, swap_rb));
    return BV_OK;
}
