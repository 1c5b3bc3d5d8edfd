// mm_tape.cu -- host side of the taped moment match + the reverse sweep that consumes the tape.
// See mm_tape.cuh for the idea.  Reference: the policy gradient of pilco/models/pilco.py:47-50, 84-90 (TensorFlow
// autodiff through mgpr.py:91-149); numpy statement of exactly these formulas: oracle/staged.py:mm_backward_tape.
#include "mm_tape.cuh"

#include <stdlib.h>

template <int KS>
static int launch_tape_tile(const MMParams& p, cudaStream_t st) {
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};      // function attributes are per device
    static int variant = 1;        // 0: <=128 regs; 1: <=96 regs, 2 CTAs/SM; 2: <=80 regs, 2 CTAs/SM (padded smem); 3: <=80 regs, 3 CTAs/SM
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {
        const char* e = getenv("PILCO_TAPE_VARIANT");           // tuning switch
        if (e && e[0] >= '0' && e[0] <= '3') variant = e[0] - '0';
        const int big = (int)mm_tape_smem_bytes(TAPE_MAX_NP, 20);
        if (cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 304>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_tile_kernel<KS, 352>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess)
            return PILCO_ERR_LAUNCH;
        configured = true;
    }
    if (p.L.np > TAPE_MAX_NP) return PILCO_ERR_UNSUPPORTED;
    size_t smem = mm_tape_smem_bytes(p.L.np, p.L.ldz);
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

// -------------------------------------------------------------------------------------------------
// reverse sweep from the tape: one CTA (4 warps) per task
//   task <  E : mean / V block of output a   (per-centre weights recomputed: n exps, W-form of mgpr.py:103-118)
//   task >= E : covariance block of the unordered pair q = task - E, from hr, hc, HZ on the tape
// Both reduce to weighted moment sums over the centres; with zx_n = [zeta_n, 1] (the 1 at index D) they are the three
// products   zx' diag(u) zx,   zx' diag(v) zx,   zx' HZ   (K = n), evaluated with fp64 DMMA m8n8k4: the k-steps (4
// centres each) are dealt round-robin to the warps, one value zx[row t][col g] per lane serves as A fragment
// (A[m=col][k=row]) AND, scaled by u / v, as B fragment (B[k=row][n=col]).  A1 = block[:D,:D], y1 = block[:D, D],
// sum u = block[D, D]; the symmetric blocks skip their lower tiles.
// -------------------------------------------------------------------------------------------------
#define TB_THREADS 128
#define TB_WARPS 4

template <int DP>
__global__ void __launch_bounds__(TB_THREADS, 6) mm_tape_bfinish_kernel(MMTapeBwd bp) {
    PDL_ENTRY();
    extern __shared__ __align__(16) double tb_dyn[];          // [2][np]: per-centre weights u, v
    const pilco_gp_model& gp = bp.gp;
    const MMTapeL& TL = bp.TL;
    const int n = gp.n, D = gp.D, E = gp.E, np = TL.np;
    const int r = blockIdx.y, task = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int TX = (DP + 8) / 8;                            // 8-wide tiles covering [zeta, 1]: D + 1 <= 8 TX
    constexpr int DX = 8 * TX;
    constexpr int NSYMT = TX * (TX + 1) / 2;                    // upper tiles of a symmetric block
    constexpr int NACC = 2 * (2 * NSYMT + TX * TX);             // accumulator doubles per lane

    __shared__ double sW[MAXD * SLD], sCm[MAXD * SLD], sT[MAXD * SLD], sX[MAXD * SLD];
    __shared__ double sinvd[MAXD], spa[MAXD], spb[MAXD], sgv[MAXD], swgv[MAXD], sm[MAXD];
    __shared__ double sA1[DX * DX], sA2[DX * DX], sA3[DX * DX], sscal[4];
    double* su = tb_dyn;
    double* sv = tb_dyn + np;
    double* sRed = tb_dyn + 2 * (size_t)np;                    // [(TB_WARPS - 1)][NACC][32] cross-warp reduction

    const double* X = gp.X + (size_t)r * gp.X_bs;
    const double* ell = gp.ell + (size_t)r * gp.ell_bs;
    const double* sf2 = gp.sf2 + (size_t)r * gp.sf2_bs;
    const double* beta = gp.beta + (size_t)r * gp.beta_bs;
    const double* mr = bp.m + (size_t)r * bp.m_rs;
    const double* sr = bp.s + (size_t)r * bp.s_rs;
    const double* gS = bp.gS + (size_t)r * E * E;
    const double* tpr = bp.tape + (size_t)r * TL.per_r;
    double* part = bp.part + ((size_t)r * (E + TL.P) + task) * (MAXD + (size_t)D * D);
    double* Tm = part;
    double* Ts = part + MAXD;
    if (tid < MAXD) sm[tid] = tid < D ? mr[tid] : 0.0;

    const bool is_out = task < E;
    int a = task, b = task, q = 0;
    const double* HZg = nullptr;
    const double* hrg = nullptr;
    const double* hcg = nullptr;
    if (is_out) {
        // ---- output task: W_a = (s + Lambda_a^2)^-1, c_a, per-centre weights u_n = gw_n w_n, v_n = w_n ----
        if (tid < DP) {
            const double l = tid < D ? ell[a * D + tid] : 1.0;
            spa[tid] = l * l;
            sgv[tid] = tid < D ? bp.gV[((size_t)r * D + tid) * E + a] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < DP * DP; e += blockDim.x) {
            const int i = e / DP, j = e % DP;
            const double sij = (i < D && j < D) ? 0.5 * (sr[i * D + j] + sr[j * D + i]) : 0.0;
            sT[i * SLD + j] = sij + (i == j ? spa[i] : 0.0);
            sW[i * SLD + j] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            chol_warp(sT, sinvd, DP, lane);
            chol_solve_warp(sT, sinvd, sW, DP, DP, lane);
            if (lane == 0) {
                double ld = chol_logdet(sinvd, DP), sl = 0.0;
                for (int d = 0; d < D; ++d) sl += log(spa[d]);
                sscal[0] = exp(log(sf2[a]) + 0.5 * (sl - ld));           // c_a
                double gmt = bp.gM[(size_t)r * E + a];                   // gMtot[a] = gM[a] - sum_b (gS[a,b]+gS[b,a]) M_b
                for (int bb = 0; bb < E; ++bb) gmt -= (gS[a * E + bb] + gS[bb * E + a]) * bp.Mfwd[(size_t)r * E + bb];
                sscal[1] = gmt;
            }
        }
        __syncthreads();
        if (tid < DP) {                       // W gV_a
            double v = 0.0;
            for (int j = 0; j < DP; ++j) v = fma(sW[tid * SLD + j], sgv[j], v);
            swgv[tid] = v;
        }
        __syncthreads();
        const double ca = sscal[0], gmt = sscal[1];
        for (int nn = tid; nn < np; nn += blockDim.x) {
            double u = 0.0, v = 0.0;
            if (nn < n) {
                double z[DP];
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = d < D ? X[(size_t)nn * D + d] - sm[d] : 0.0;
                double e = 0.0, tg = 0.0;
#pragma unroll 1
                for (int i = 0; i < D; ++i) {                            // (rolled: keeps W out of the registers)
                    double ti = 0.0;
#pragma unroll
                    for (int j = 0; j < DP; ++j) ti = fma(sW[i * SLD + j], z[j], ti);
                    const double zi = X[(size_t)nn * D + i] - sm[i];
                    e = fma(zi, ti, e);
                    tg = fma(zi, swgv[i], tg);                           // t_n . gV_a = zeta_n . (W gV_a)
                }
                v = beta[(size_t)a * n + nn] * exp(-0.5 * e) * ca;       // w_n
                u = (gmt + tg) * v;                                      // gw_n w_n
            }
            su[nn] = u; sv[nn] = v;
        }
    } else {
        // ---- pair task: Q, C from the tape; weights u = hr, v = hc (summed over the row splits) ----
        q = task - E;
        pair_decode(q, a, b);
        if (tid < DP) {
            const double la = tid < D ? ell[a * D + tid] : 1.0, lb = tid < D ? ell[b * D + tid] : 1.0;
            spa[tid] = tid < D ? 1.0 / (la * la) : 0.0;
            spb[tid] = tid < D ? 1.0 / (lb * lb) : 0.0;
        }
        const double* Qg = tpr + TL.Q + (size_t)q * D * D;
        const double* Cg = tpr + TL.C + (size_t)q * D * D;
        for (int e = tid; e < DP * DP; e += blockDim.x) {
            const int i = e / DP, j = e % DP;
            const bool in = i < D && j < D;
            sW[i * SLD + j] = in ? Qg[i * D + j] : 0.0;
            sCm[i * SLD + j] = in ? Cg[i * D + j] : 0.0;
        }
        hrg = tpr + TL.hr + (size_t)q * np;                     // weights u = hr, v = sum of the row splits of hc:
        hcg = tpr + TL.hc + (size_t)q * TL.cs * np;             // read straight from the tape inside the k-loop
        HZg = tpr + TL.HZ + (size_t)q * np * TL.ldh;
    }
    __syncthreads();

    // ---- weighted moment sums by DMMA ----
    double cu[NSYMT][2], cv[NSYMT][2], ch[TX * TX][2];
#pragma unroll
    for (int i = 0; i < NSYMT; ++i) { cu[i][0] = cu[i][1] = cv[i][0] = cv[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < TX * TX; ++i) { ch[i][0] = ch[i][1] = 0.0; }
    // TBU k-steps per trip: all their operands are loaded first (the tape was written a whole rollout ago and comes
    // from DRAM -- the loads of a trip are independent, so their latencies overlap), then the DMMAs run
    constexpr int TBU = 2;
    const int nks = (n + 3) >> 2;
    for (int ks0 = warp; ks0 < nks; ks0 += TB_WARPS * TBU) {
        double zx[TBU][TX], bh[TBU][TX], uu[TBU], vv[TBU];
#pragma unroll
        for (int k = 0; k < TBU; ++k) {
            const int row = 4 * (ks0 + k * TB_WARPS) + t;
            const bool live = row < n;                            // (also false for k-steps beyond nks)
            double uk = 0.0, vk = 0.0;
            if (live) {
                if (is_out) { uk = su[row]; vk = sv[row]; }
                else { uk = hrg[row]; for (int c2 = 0; c2 < TL.cs; ++c2) vk += hcg[(size_t)c2 * np + row]; }
            }
            uu[k] = uk; vv[k] = vk;
#pragma unroll
            for (int tl = 0; tl < TX; ++tl) {
                const int c = g + 8 * tl;
                double z = 0.0, h = 0.0;
                if (live) {
                    if (c < D) { z = X[(size_t)row * D + c]; if (HZg) h = HZg[(size_t)row * TL.ldh + c]; }
                    else if (c == D) z = 1.0;
                }
                zx[k][tl] = z; bh[k][tl] = h;
            }
        }
#pragma unroll
        for (int k = 0; k < TBU; ++k) {
            double bu[TX], bv[TX];
#pragma unroll
            for (int tl = 0; tl < TX; ++tl) {
                const int c = g + 8 * tl;
                if (c < D) zx[k][tl] -= sm[c];                    // (a dead row has z = 0 and u = v = h = 0: no contribution)
                bu[tl] = uu[k] * zx[k][tl]; bv[tl] = vv[k] * zx[k][tl];
            }
            int si = 0;
#pragma unroll
            for (int mt = 0; mt < TX; ++mt)
#pragma unroll
                for (int nt = 0; nt < TX; ++nt) {
                    if (nt >= mt) {
                        dmma884(cu[si][0], cu[si][1], zx[k][mt], bu[nt]);
                        dmma884(cv[si][0], cv[si][1], zx[k][mt], bv[nt]);
                        ++si;
                    }
                    if (!is_out) dmma884(ch[mt * TX + nt][0], ch[mt * TX + nt][1], zx[k][mt], bh[k][nt]);
                }
        }
    }
    // cross-warp reduction (fixed order), then warp 0 scatters the C fragments into square matrices
    if (warp > 0) {
        double* dst = sRed + ((size_t)(warp - 1) * NACC) * 32 + lane;
        int k = 0;
#pragma unroll
        for (int i = 0; i < NSYMT; ++i) { dst[32 * k++] = cu[i][0]; dst[32 * k++] = cu[i][1]; dst[32 * k++] = cv[i][0]; dst[32 * k++] = cv[i][1]; }
#pragma unroll
        for (int i = 0; i < TX * TX; ++i) { dst[32 * k++] = ch[i][0]; dst[32 * k++] = ch[i][1]; }
    }
    __syncthreads();
    if (warp == 0) {
        for (int w = 0; w < TB_WARPS - 1; ++w) {
            const double* src = sRed + ((size_t)w * NACC) * 32 + lane;
            int k = 0;
#pragma unroll
            for (int i = 0; i < NSYMT; ++i) { cu[i][0] += src[32 * k++]; cu[i][1] += src[32 * k++]; cv[i][0] += src[32 * k++]; cv[i][1] += src[32 * k++]; }
#pragma unroll
            for (int i = 0; i < TX * TX; ++i) { ch[i][0] += src[32 * k++]; ch[i][1] += src[32 * k++]; }
        }
        int si = 0;
#pragma unroll
        for (int mt = 0; mt < TX; ++mt)
#pragma unroll
            for (int nt = 0; nt < TX; ++nt) {
                const int i = 8 * mt + g, j = 8 * nt + 2 * t;
                if (nt >= mt) {
                    sA1[i * DX + j] = cu[si][0]; sA1[i * DX + j + 1] = cu[si][1];
                    sA2[i * DX + j] = cv[si][0]; sA2[i * DX + j + 1] = cv[si][1];
                    if (nt > mt) {
                        sA1[j * DX + i] = cu[si][0]; sA1[(j + 1) * DX + i] = cu[si][1];
                        sA2[j * DX + i] = cv[si][0]; sA2[(j + 1) * DX + i] = cv[si][1];
                    }
                    ++si;
                }
                sA3[i * DX + j] = ch[mt * TX + nt][0]; sA3[i * DX + j + 1] = ch[mt * TX + nt][1];
            }
    }
    __syncthreads();
    // A1 = sA1[:D,:D] etc.; y1[i] = sA1[i][D], y2[i] = sA2[i][D]; sum u = sA1[D][D], sum v = sA2[D][D]
    const double sum_u = sA1[D * DX + D], sum_v = sA2[D * DX + D];

    if (is_out) {
        // gW = -0.5 A1 + sym(gV y2');  gA = -W gW W - 0.5 glogc W;  sum gzeta = -W y1 + (sum w) W gV
        const double glogc = sum_u;
        for (int e = tid; e < DP * DP; e += blockDim.x) {
            const int i = e / DP, j = e % DP;
            sT[i * SLD + j] = (i < D && j < D) ? -0.5 * sA1[i * DX + j] + 0.5 * (sgv[i] * sA2[j * DX + D] + sgv[j] * sA2[i * DX + D]) : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < DP * DP; e += blockDim.x) {           // sCm = W gW
            const int i = e / DP, j = e % DP;
            double v = 0.0;
            for (int k = 0; k < DP; ++k) v = fma(sW[i * SLD + k], sT[k * SLD + j], v);
            sCm[i * SLD + j] = v;
        }
        __syncthreads();
        for (int e = tid; e < D * D; e += blockDim.x) {
            const int i = e / D, j = e % D;
            double v = 0.0;
            for (int k = 0; k < DP; ++k) v = fma(sCm[i * SLD + k], sW[k * SLD + j], v);
            Ts[e] = -v - 0.5 * glogc * sW[i * SLD + j];
        }
        if (tid < D) {
            double v = 0.0;
            for (int j = 0; j < D; ++j) v = fma(sW[tid * SLD + j], sA1[j * DX + D], v);
            Tm[tid] = v - sum_v * swgv[tid];                        // -(sum_n gzeta_n)
        }
        return;
    }

    // pair task
    const double gw = (a == b) ? gS[a * E + a] : gS[a * E + b] + gS[b * E + a];
    const double glogR = -0.5 * gw * sum_u;
    for (int e = tid; e < DP * DP; e += blockDim.x) {               // sT = gQ / (delta_i delta_j)
        const int i = e / DP, j = e % DP;
        double v = 0.0;
        if (i < D && j < D) {
            const double gq = gw * (spa[i] * spa[j] * sA1[i * DX + j] + spb[i] * spb[j] * sA2[i * DX + j]
                                    + spa[i] * spb[j] * sA3[i * DX + j] + spa[j] * spb[i] * sA3[j * DX + i]);
            v = gq / ((spa[i] + spb[i]) * (spa[j] + spb[j]));
        }
        sT[i * SLD + j] = v;
    }
    __syncthreads();
    for (int e = tid; e < DP * DP; e += blockDim.x) {               // sX = C (gQ/dd)
        const int i = e / DP, j = e % DP;
        double v = 0.0;
        for (int k = 0; k < D; ++k) v = fma(sCm[i * SLD + k], sT[k * SLD + j], v);
        sX[i * SLD + j] = v;
    }
    __syncthreads();
    for (int e = tid; e < D * D; e += blockDim.x) {                 // gs_pair = 0.5 C (gQ/dd) C + glogR C
        const int i = e / D, j = e % D;
        double v = 0.0;
        for (int k = 0; k < D; ++k) v = fma(sX[i * SLD + k], sCm[k * SLD + j], v);
        Ts[e] = 0.5 * v + glogR * sCm[i * SLD + j];
    }
    if (tid < D) {                                                  // -(sum gzeta) = g (u - 2 delta o Q u),  u = p_a y1 + p_b y2
        double qu = 0.0;
        for (int j = 0; j < D; ++j) qu = fma(sW[tid * SLD + j], spa[j] * sA1[j * DX + D] + spb[j] * sA2[j * DX + D], qu);
        const double u = spa[tid] * sA1[tid * DX + D] + spb[tid] * sA2[tid * DX + D];
        Tm[tid] = gw * (u - 2.0 * (spa[tid] + spb[tid]) * qu);
    }
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
    // dynamic shared memory: per-centre weights [2][np] + the cross-warp reduction buffer [3][NACC][32]
    const int dp = 4 * ksteps_of(bp.gp.D), tx = (dp + 8) / 8;
    const int nacc = 2 * (tx * (tx + 1) + tx * tx);
    const size_t smem = ((size_t)2 * bp.TL.np + (size_t)(TB_WARPS - 1) * nacc * 32) * sizeof(double);
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {                                          // large n: static + dynamic shared memory > 48 KB
        const int big = (2 * TAPE_MAX_NP + (TB_WARPS - 1) * 42 * 32) * (int)sizeof(double);
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
    if (n < 1 || D < 1 || D > MAXD || E < 1 || E > MAXE || R < 1 || pad64(n) > TAPE_MAX_NP) return 0;
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
    if (pad64(gp->n) > TAPE_MAX_NP) return PILCO_ERR_UNSUPPORTED;
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
    if (pad64(gp->n) > TAPE_MAX_NP) return PILCO_ERR_UNSUPPORTED;
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
