/* C99 host of the drop-in boundary: include/xr_b200.h must compile as plain C and libxr_b200.so must link and load without
 * Python or torch.  Built and run by tests/test_abi_cpu.py.  Exit code 0 = as expected for this machine:
 *   without a usable sm_100 device xr_ctx_create returns XR_ERR_NO_DEVICE and leaves a message (no CPU fallback);
 *   with one, a context is created, reports an sm_100 device, runs a 2x2x2 xr_gemm_scatter and a legacy scalar, and is destroyed. */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "xr_b200.h"

int main(void) {
    const char* version = xr_version();
    xr_ctx* ctx = NULL;
    int rc;
    if (!version || !strstr(version, "sm_100a")) {
        fprintf(stderr, "unexpected version string\n");
        return 2;
    }
    rc = xr_ctx_create(0, NULL, 1, &ctx);
    if (rc == XR_ERR_NO_DEVICE) {
        const char* msg = xr_last_error();
        double nan_result = monomer_1e(2, NULL, NULL);      /* legacy scalar ABI: NaN, never a CPU result */
        printf("no device: %s\n", msg ? msg : "(null)");
        return (msg && msg[0] && ctx == NULL && isnan(nan_result)) ? 0 : 3;
    }
    if (rc != XR_OK) {
        fprintf(stderr, "xr_ctx_create: %d %s\n", rc, xr_last_error());
        return 4;
    }
    {
        int sm = 0, major = 0, minor = 0;
        size_t free_b = 0, total_b = 0;
        const double a[4] = {1, 2, 3, 4}, b[4] = {5, 6, 7, 8};
        double c[4] = {0, 0, 0, 0};
        void *da = NULL, *db = NULL, *dc = NULL;
        if (xr_device_info(ctx, &sm, &major, &minor, &free_b, &total_b) != XR_OK || major != 10) return 5;
        if (xr_malloc(ctx, sizeof a, &da) || xr_malloc(ctx, sizeof b, &db) || xr_malloc(ctx, sizeof c, &dc)) return 6;
        if (xr_upload(ctx, da, a, sizeof a) || xr_upload(ctx, db, b, sizeof b)) return 7;
        /* C[m*2 + n] = sum_k A[m*2 + k] * B[n*2 + k] */
        if (xr_gemm_scatter(ctx, 2, 2, 2, 1.0, (const double*)da, 2, (const double*)db, 2, (double*)dc, NULL, 2, NULL, 0)) return 8;
        if (xr_download(ctx, c, dc, sizeof c) || xr_sync(ctx)) return 9;
        if (c[0] != 17.0 || c[1] != 23.0 || c[2] != 39.0 || c[3] != 53.0) return 10;
        xr_free(ctx, da);
        xr_free(ctx, db);
        xr_free(ctx, dc);
        printf("device sm_%d%d, %d SMs: gemm ok\n", major, minor, sm);
        {   /* legacy scalar ABI with HOST buffers (H_contractions.c:48): sum_pq Rca[p,q] h[p,q] = 5+12+21+32 */
            double rca[4] = {1, 2, 3, 4}, h[4] = {5, 6, 7, 8};
            if (monomer_1e(2, rca, h) != 70.0) return 12;
            printf("legacy scalar ok\n");
        }
    }
    return xr_ctx_destroy(ctx) == XR_OK ? 0 : 11;
}
