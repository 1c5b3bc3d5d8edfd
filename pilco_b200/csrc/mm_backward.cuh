// mm_backward.cuh -- workspace layout and parameter block of the moment-match VJP (see mm_backward.cu).
#pragma once
#include "mm_kernels.cuh"

struct MMBws {                      // extra workspace arrays (offsets in doubles, per restart)
    MMWs F;                         // forward-style arrays for ORDERED pairs
    size_t oQ, oC, oLd, rowout, Tm, Ts, Tpa, Tpb, Tsc, Tz, Tb, cnt, per_r;
    int ldr, ntask, P2;
};

static inline __host__ __device__ MMBws mm_bws_layout(int n, int D, int E, int need_param) {
    MMBws B;
    B.F = mm_ws_layout(n, D, E, true);
    const int DP = 4 * ksteps_of(D);
    B.P2 = E * E; B.ntask = E + E * E; B.ldr = 2 * (DP + 1);
    size_t o = B.F.per_r;
    auto take = [&](size_t len) { size_t at = o; o += (len + 1) & ~(size_t)1; return at; };
    B.oQ = take((size_t)B.P2 * D * D);
    B.oC = take((size_t)B.P2 * D * D);
    B.oLd = take(B.P2);
    B.rowout = take((size_t)B.P2 * B.F.np * B.ldr);
    B.Tm = take((size_t)B.ntask * MAXD);
    B.Ts = take((size_t)B.ntask * D * D);
    B.Tpa = take((size_t)B.ntask * MAXD);
    B.Tpb = take((size_t)B.ntask * MAXD);
    B.Tsc = take((size_t)B.ntask * 2);
    B.Tz = need_param ? take((size_t)B.ntask * B.F.np * MAXD) : o;
    B.Tb = need_param ? take((size_t)B.ntask * B.F.np) : o;
    B.cnt = take(2);                // arrival counter of the finish tasks (last_cta_arrives); zero before first use
    B.per_r = o;
    return B;
}

struct MMBwdParams {
    MMParams f;                     // model, inputs m/s, forward outputs M (f.M) are read; f.ws = workspace
    MMBws B;
    const double* gM; const double* gS; const double* gV;     // upstream cotangents [R,E],[R,E,E],[R,D,E]
    double* gm; double* gs;         // [R,D] (row stride gm_rs), [R,D,D] (stride gs_rs)
    long long gm_rs, gs_rs;
    double* gX; double* gbeta; double* gell;                  // [R,n,D],[R,E,n],[R,E,D] or NULL
    int need_param, accumulate;
};


MMBwdParams mm_bwd_params(const pilco_gp_model* gp, int R, const double* m, long long m_rs, const double* s, long long s_rs,
                          const double* Mfwd, const double* gM, const double* gS, const double* gV,
                          double* gm, long long gm_rs, double* gs, long long gs_rs,
                          double* gX, double* gbeta, double* gell, int accumulate, double* ws);
int mm_backward_launch(MMBwdParams bp, cudaStream_t st);
// zero the arrival counters of a backward workspace (once before the first mm_backward_launch on it)
static inline void mm_bwd_zero_counters(const MMBws& B, double* ws, int R, cudaStream_t st) {
    cudaMemset2DAsync(ws + B.cnt, B.per_r * sizeof(double), 0, 2 * sizeof(double), (size_t)R, st);
}
