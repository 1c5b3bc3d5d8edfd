/* A plain C translation unit that binds libpilco_b200.so through include/pilco_b200.h only (no C++, no torch):
 * the CPU suite compiles and links it and runs the GPU-free part (version, sizing, argument validation);
 * `run_moment_match` shows the call sequence a C host would make with its own cudaMalloc'ed buffers. */
#include <stdio.h>
#include <string.h>
#include "../../include/pilco_b200.h"

/* device pointers owned by the caller; returns the library status */
int run_moment_match(int n, int D, int E, const double* dX, const double* dY, const double* dell, const double* dsf2,
                     const double* dsn2, double* diK, double* dbeta, void* dws_fact, size_t ws_fact_bytes,
                     const double* dm, const double* ds, double* dM, double* dS, double* dV,
                     void* dws_mm, size_t ws_mm_bytes, pilco_stream_t stream) {
    const int ldk = pilco_pad_n(n);
    int rc = pilco_gp_factorize(n, D, E, 1, dX, 0, dY, 0, dell, 0, dsf2, 0, dsn2, 0, diK, ldk, dbeta, NULL,
                                dws_fact, ws_fact_bytes, stream);              /* mgpr.py:81-89 */
    if (rc) return rc;
    pilco_gp_model gp;
    memset(&gp, 0, sizeof gp);
    gp.n = n; gp.D = D; gp.E = E; gp.mode = 0;
    gp.X = dX; gp.ell = dell; gp.sf2 = dsf2; gp.beta = dbeta; gp.iK = diK; gp.ldk = ldk;
    return pilco_mm_forward(&gp, 1, dm, ds, dM, dS, dV, NULL, dws_mm, ws_mm_bytes, stream);   /* mgpr.py:91-149 */
}

int main(void) {
    if (pilco_version() != PILCO_ABI_VERSION) { printf("abi mismatch\n"); return 1; }
    if (pilco_pad_n(300) != 320) return 2;
    const size_t a = pilco_mm_workspace_bytes(300, 12, 10, 32), b = pilco_gp_factorize_workspace_bytes(300, 10, 1);
    if (a == 0 || b == 0) return 3;
    if (pilco_mm_workspace_bytes(300, PILCO_MAX_D + 1, 10, 1) != 0) return 4;
    /* invalid arguments are rejected on the host, before anything is enqueued */
    if (pilco_mm_forward(NULL, 1, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) >= 0) return 5;
    printf("abi %d mm_ws %zu fact_ws %zu status(-3)=%s\n", pilco_version(), a, b, pilco_status_string(-3));
    return 0;
}
