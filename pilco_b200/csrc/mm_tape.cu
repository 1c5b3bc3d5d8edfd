// mm_tape.cu -- host side of the taped moment match + the reverse sweep that consumes the tape.
// See mm_tape.cuh for the idea.  Reference: the policy gradient of pilco/models/pilco.py:47-50, 84-90 (TensorFlow
// autodiff through mgpr.py:91-149); numpy statement of exactly these formulas: oracle/staged.py:mm_backward_tape.
#include "mm_tape.cuh"

#include <stdlib.h>

template <int KS>
static int launch_tape_tile(const MMParams& p, cudaStream_t st) {
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};      // function attributes are per device
    static int variant = 0;        // 0: <=128 regs (default: fastest tile pass, measured); 1: <=96 regs, 2 CTAs/SM; 2: <=80 regs, 2 CTAs/SM (padded smem); 3: <=80 regs, 3 CTAs/SM
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {
        const char* e = getenv("PILCO_TAPE_VARIANT");           // tuning switch
        if (e && e[0] >= '0' && e[0] <= '3') variant = e[0] - '0';
        const int big = PILCO_MAX_SMEM_OPTIN;
        if (cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 304>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 352>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess)
            return PILCO_ERR_LAUNCH;
        configured = true;
    }
    size_t smem = mm_tape_smem_bytes(p.L.np, p.L.ldz);
    if (p.L.np > TAPE_MAX_NP || smem > PILCO_MAX_SMEM_OPTIN) return PILCO_ERR_UNSUPPORTED;
    const dim3 grid(p.TL.cs, p.L.P, p.R);
    const bool hi = pilco_small_grid(grid);
    if (variant == 0) launch_pri(hi, mm_tape_tile_kernel<KS, 256>, grid, dim3(256), smem, st, p);
    else if (variant == 1) launch_pri(hi, mm_tape_tile_kernel<KS, 304>, grid, dim3(256), smem, st, p);
    else {
        if (variant == 2 && smem < 80 * 1024) smem = 80 * 1024;    // 2 CTAs per SM
        launch_pri(hi, mm_tape_tile_kernel<KS, 352>, grid, dim3(256), smem, st, p);
    }
    return PILCO_OK;
}

int mm_tape_tile_launch(const MMParams& p, cudaStream_t st) {
    switch (ksteps_of(p.gp.D)) {
        case 1: return launch_tape_tile<1>(p, st);
        case 2: return launch_tape_tile<2>(p, st);
        case 3: return launch_tape_tile<3>(p, st);
        default: return launch_tape_tile<4>(p, st);
    }
}

template <int DP>
__global__ void __launch_bounds__(TB_THREADS, 3) mm_tape_bfinish_kernel(MMTapeBwd bp) {
    PDL_ENTRY();
    extern __shared__ __align__(16) double tb_dyn[];
    mm_tape_bfinish_task<DP>(bp, blockIdx.y, blockIdx.x, tb_dyn);
}

// sum the task partials -> gm, gs
__global__ void __launch_bounds__(128) mm_tape_breduce_kernel(MMTapeBwd bp) {
    PDL_ENTRY();
    const int r = blockIdx.x;
    mm_tape_reduce_device(bp.part + (size_t)r * mm_tape_bwd_part_doubles(bp.gp.D, bp.gp.E), bp.gp.E + bp.TL.P, bp.gp.D,
                          bp.gm + (size_t)r * bp.gm_rs, bp.gs + (size_t)r * bp.gs_rs, bp.accumulate);
}

int mm_tape_backward_launch(const MMTapeBwd& bp, cudaStream_t st, bool with_reduce) {
    const int E = bp.gp.E, R = bp.R;
    dim3 gf(E + bp.TL.P, R);
    const size_t smem = mm_tape_bfinish_smem_bytes(bp.TL.np, bp.gp.D);
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {                                          // large n: static + dynamic shared memory > 48 KB
        const int big = 196 * 1024;
        if (cudaFuncSetAttribute(mm_tape_bfinish_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess)
            return PILCO_ERR_LAUNCH;
        configured = true;
    }
    switch (ksteps_of(bp.gp.D)) {
        case 1: launch_hi(mm_tape_bfinish_kernel<4>, dim3(gf), dim3(TB_THREADS), smem, st, bp); break;
        case 2: launch_hi(mm_tape_bfinish_kernel<8>, dim3(gf), dim3(TB_THREADS), smem, st, bp); break;
        case 3: launch_hi(mm_tape_bfinish_kernel<12>, dim3(gf), dim3(TB_THREADS), smem, st, bp); break;
        default: launch_hi(mm_tape_bfinish_kernel<16>, dim3(gf), dim3(TB_THREADS), smem, st, bp); break;
    }
    CUDA_LAUNCH_CHECK();
    if (with_reduce) {
        launch_hi(mm_tape_breduce_kernel, dim3(R), dim3(128), 0, st, bp);
        CUDA_LAUNCH_CHECK();
    }
    return PILCO_OK;
}

extern "C" {

size_t pilco_mm_tape_bytes(int n, int D, int E, int R) {
    if (n < 1 || D < 1 || D > MAXD || E < 1 || E > MAXE || R < 1 || !mm_tape_supported(n, D)) return 0;
    return mm_tape_layout(n, D, E, R).per_r * (size_t)R * sizeof(double);
}

size_t pilco_mm_tape_bwd_workspace_bytes(int D, int E, int R) {
    if (D < 1 || D > MAXD || E < 1 || E > MAXE || R < 1) return 0;
    return mm_tape_bwd_part_doubles(D, E) * (size_t)R * sizeof(double);
}

int pilco_mm_forward_taped(const pilco_gp_model* gp, int R, const double* m, const double* s,
                           double* M, double* S, double* V, int* info,
                           void* ws, size_t ws_bytes, void* tape, size_t tape_bytes, pilco_stream_t stream) {
    int rc = mm_check_model(gp);
    if (rc) return rc;
    if (!m || !s || !M || !S || !V || !ws || !tape) return PILCO_ERR_NULL;
    if (R < 1) return PILCO_ERR_DIM;
    if (!mm_tape_supported(gp->n, gp->D)) return PILCO_ERR_UNSUPPORTED;
    if (ws_bytes < pilco_mm_workspace_bytes(gp->n, gp->D, gp->E, R)) return PILCO_ERR_WORKSPACE;
    if (tape_bytes < pilco_mm_tape_bytes(gp->n, gp->D, gp->E, R)) return PILCO_ERR_WORKSPACE;
    if ((((uintptr_t)ws) | ((uintptr_t)tape)) & 15) return PILCO_ERR_ALIGN;
    MMParams p;
    p.gp = *gp; p.R = R; p.m = m; p.s = s; p.m_rs = gp->D; p.s_rs = (long long)gp->D * gp->D; p.M = M; p.S = S; p.V = V; p.info = info;
    p.ws = (double*)ws; p.L = mm_ws_layout(gp->n, gp->D, gp->E); p.bwd = 0; p.oQ = p.oC = p.oLd = 0;
    p.tape = (double*)tape; p.TL = mm_tape_layout(gp->n, gp->D, gp->E, R);
    return mm_forward_launch(p, (cudaStream_t)stream, true);
}

int pilco_mm_backward_taped(const pilco_gp_model* gp, int R, const double* m, const double* s, const double* M,
                            const double* gM, const double* gS, const double* gV,
                            const void* tape, size_t tape_bytes, double* gm, double* gs,
                            void* ws, size_t ws_bytes, pilco_stream_t stream) {
    int rc = mm_check_model(gp);
    if (rc) return rc;
    if (!m || !s || !M || !gM || !gS || !gV || !tape || !gm || !gs || !ws) return PILCO_ERR_NULL;
    if (R < 1) return PILCO_ERR_DIM;
    if (!mm_tape_supported(gp->n, gp->D)) return PILCO_ERR_UNSUPPORTED;
    if (tape_bytes < pilco_mm_tape_bytes(gp->n, gp->D, gp->E, R)) return PILCO_ERR_WORKSPACE;
    if (ws_bytes < pilco_mm_tape_bwd_workspace_bytes(gp->D, gp->E, R)) return PILCO_ERR_WORKSPACE;
    if ((((uintptr_t)ws) | ((uintptr_t)tape)) & 15) return PILCO_ERR_ALIGN;
    MMTapeBwd bp;
    bp.gp = *gp; bp.R = R; bp.m = m; bp.s = s; bp.m_rs = gp->D; bp.s_rs = (long long)gp->D * gp->D;
    bp.Mfwd = M; bp.gM = gM; bp.gS = gS; bp.gV = gV;
    bp.tape = (const double*)tape; bp.TL = mm_tape_layout(gp->n, gp->D, gp->E, R);
    bp.part = (double*)ws; bp.gm = gm; bp.gs = gs; bp.gm_rs = gp->D; bp.gs_rs = (long long)gp->D * gp->D;
    bp.accumulate = 0;
    return mm_tape_backward_launch(bp, (cudaStream_t)stream, true);
}

}  // extern "C"
