/* A plain C99 consumer of the drop-in boundary: include/camcal_b200.h + libcamcal_b200.so, nothing else.
 * What a non-Python, non-Julia caller sees.  Exit code 0 = every check passed; prints "no device" and
 * still exits 0 when the box has no GPU (the library must say CC_ERR_NO_DEVICE, not fall back).
 * Built and run by tests/test_abi.py (CPU) and tests/test_gpu_parity.py (GPU). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "camcal_b200.h"

#define CHECK(cond)                                                                     \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            fprintf(stderr, "FAILED %s (line %d): %s\n", #cond, __LINE__, cc_last_error_string()); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

int main(void)
{
    cc_ctx *ctx = NULL;
    int rc = cc_ctx_create(0, &ctx);
    /* host-side helpers work anywhere */
    {
        double rows[6] = {10, 20, 30, 10, 20, 30}, cols[6] = {5, 5, 5, 15, 15, 15}, ratio = 0;
        int64_t axs[2];
        CHECK(cc_get_ratio(rows, cols, 3, 2, 2.0, &ratio) == CC_OK);
        CHECK(fabs(ratio - 5.0) < 1e-12);                 /* mean step 10 px per 2 units */
        CHECK(cc_get_axes(ratio, 2.0, 3, 2, 100, 80, axs) == CC_OK);
        CHECK(axs[0] == -40 && axs[1] == -35);            /* round((20 - 100) / 2), round((10 - 80) / 2) */
        CHECK(cc_get_ratio(NULL, cols, 3, 2, 2.0, &ratio) == CC_ERR_INVALID_ARG);
    }
    if (rc == CC_ERR_NO_DEVICE) {
        printf("no device\n");
        return 0;
    }
    CHECK(rc == CC_OK && ctx != NULL);

    /* pinhole looking straight at the board: f = 128, c = (50, 60), k = 0, R = I, t = (0, 0, 16):
     * powers of two, so every intermediate of the chain is exact */
    cc_intr intr = {128.0, 128.0, 50.0, 60.0, 0.0, 1.0};
    cc_view view = {{0.0, 0.0, 0.0}, {0.0, 0.0, 16.0}};
    enum { N = 5 };
    double x[N] = {0, 1, -2, 3.5, 0.25}, y[N] = {0, -1, 2, 0.5, -4}, z[N] = {0, 0, 0, 0, 0};
    double row[N], col[N], bx[N], by[N], bz[N];
    CHECK(cc_world2img_f64_host(ctx, &intr, &view, x, y, z, row, col, N) == CC_OK);
    for (int i = 0; i < N; ++i) {
        CHECK(fabs(row[i] - (50.0 + 8.0 * x[i])) < 1e-12);
        CHECK(fabs(col[i] - (60.0 + 8.0 * y[i])) < 1e-12);
    }
    CHECK(cc_img2world_f64_host(ctx, &intr, &view, row, col, bx, by, bz, N) == CC_OK);
    for (int i = 0; i < N; ++i) CHECK(fabs(bx[i] - x[i]) < 1e-12 && fabs(by[i] - y[i]) < 1e-12 && fabs(bz[i]) < 1e-12);

    /* rectification of a 16 x 4 ramp with ratio 8 and axes chosen so that output == input */
    float src[4 * 16], dst[4 * 16];
    for (int i = 0; i < 64; ++i) src[i] = (float)i;
    {
        /* world = I / 8; pixel = c + 128 * world / 16 = c + I: output index I1 samples row 50 + I1 -> axs_min = 1 - c */
        const int64_t a[2] = {1 - 50, 1 - 60};
        CHECK(cc_rectify_f32c1_host(ctx, &intr, &view, 8.0, a, src, dst, 16, 4, 16, 64, 1, -1.0f, CC_COORD_F64) == CC_OK);
        CHECK(memcmp(src, dst, sizeof(src)) == 0);
    }
    /* argument errors are status codes */
    CHECK(cc_img2world_f64_host(ctx, NULL, &view, row, col, bx, by, bz, N) == CC_ERR_INVALID_ARG);
    uint64_t launches = 0;
    CHECK(cc_ctx_launch_count(ctx, &launches) == CC_OK && launches >= 3);
    CHECK(cc_ctx_destroy(ctx) == CC_OK);
    printf("ok\n");
    return 0;
}
