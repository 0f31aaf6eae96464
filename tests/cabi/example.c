/* Plain C99 caller of the C ABI (no C++, no CUDA headers): what a C/C++ module of the reference tree would
 * write.  Reads a raw BGR frame, runs (1) the legacy process_frame symbol in place and (2) the fused stage
 * from host buffers, writes the results next to the input.
 *   example <frame.raw> <height> <width> <out_balanced.raw> <out_mask.raw> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200vision.h"

int main(int argc, char **argv) {
    if (argc != 6) return 2;
    const int h = atoi(argv[2]), w = atoi(argv[3]);
    const size_t n = (size_t)h * w;
    uint8_t *frame = (uint8_t *)malloc(n * 3), *bal = (uint8_t *)malloc(n * 3), *mask = (uint8_t *)malloc(n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(frame, 1, n * 3, f) != n * 3) return 3;
    fclose(f);

    /* (1) modules/color_balance.py:105-107 calls exactly this symbol */
    memcpy(bal, frame, n * 3);
    if (process_frame(bal, (size_t)h, (size_t)w, 3, true, false, true, false, true, false, 1, 1) != 0) {
        fprintf(stderr, "process_frame: %s\n", bv_last_error());
        return 4;
    }

    /* (2) modules/bins.py:13-27 as one call */
    bv_ctx *ctx = NULL;
    if (bv_create(0, &ctx) != BV_OK) {
        fprintf(stderr, "bv_create: %s\n", bv_last_error());
        return 5;
    }
    bv_stage_desc d;
    memset(&d, 0, sizeof d);
    d.cvt_code = BV_BGR2HSV;
    d.lo[0] = 10; d.lo[1] = 20; d.lo[2] = 60;
    d.hi[0] = 30; d.hi[1] = 100; d.hi[2] = 255;
    d.n_morph = 1;
    d.morph_op[0] = BV_MORPH_OPEN;
    d.morph_kw[0] = d.morph_kh[0] = 5;
    d.morph_iters[0] = 1;
    d.do_label = 1;
    bv_blob blobs[256];
    int32_t n_blobs = 0;
    if (bv_stage_host(ctx, &d, frame, 1, h, w, NULL, NULL, mask, NULL, blobs, 256, &n_blobs) != BV_OK) {
        fprintf(stderr, "bv_stage_host: %s\n", bv_last_error());
        return 6;
    }
    /* the same call split in two (bv_stage_host_submit / bv_stage_host_wait): both slots in flight, same results */
    {
        uint8_t *mask2[2];
        bv_blob blobs2[2][256];
        int32_t n2[2] = {0, 0};
        int k;
        for (k = 0; k < 2; ++k) {
            mask2[k] = (uint8_t *)malloc(n);
            if (bv_stage_host_submit(ctx, k, &d, frame, 1, h, w, NULL, NULL, mask2[k], NULL, blobs2[k], 256, &n2[k]) != BV_OK) {
                fprintf(stderr, "bv_stage_host_submit: %s\n", bv_last_error());
                return 7;
            }
        }
        for (k = 1; k >= 0; --k) {
            if (bv_stage_host_wait(ctx, k) != BV_OK || n2[k] != n_blobs || memcmp(mask2[k], mask, n) != 0 ||
                memcmp(blobs2[k], blobs, sizeof(bv_blob) * (size_t)n_blobs) != 0) {
                fprintf(stderr, "submit / wait differs from the blocking call (slot %d)\n", k);
                return 8;
            }
            free(mask2[k]);
        }
    }
    bv_destroy(ctx);

    f = fopen(argv[4], "wb"); fwrite(bal, 1, n * 3, f); fclose(f);
    f = fopen(argv[5], "wb"); fwrite(mask, 1, n, f); fclose(f);
    printf("blobs %d launches ok\n", (int)n_blobs);
    return 0;
}
