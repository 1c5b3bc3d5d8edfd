// mm_tape.cu -- host side of the taped moment match + the reverse sweep that consumes the tape.
// See mm_tape.cuh for the idea.  Reference: the policy gradient of pilco/models/pilco.py:47-50, 84-90 (TensorFlow
// autodiff through mgpr.py:91-149); numpy statement of exactly these formulas: oracle/staged.py:mm_backward_tape.
#include "mm_tape.cuh"

template <int KS>
static int launch_tape_tile(const MMParams& p, cudaStream_t st) {
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};      // function attributes are per device
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {
        const int big = (int)mm_tape_smem_bytes(TAPE_MAX_NP, 20);
        if (cudaFuncSetAttribute(mm_tape_tile_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return PILCO_ERR_LAUNCH;
        configured = true;
    }
    if (p.L.np > TAPE_MAX_NP) return PILCO_ERR_UNSUPPORTED;
    const size_t smem = mm_tape_smem_bytes(p.L.np, p.L.ldz);
    mm_tape_tile_kernel<KS><<<dim3(p.TL.cs, p.L.P, p.R), 256, smem, st>>>(p);
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
// reverse sweep from the tape: one CTA (256 threads) per task
//   task <  E : mean / V block of output a   (per-centre weights recomputed: n exps, W-form of mgpr.py:103-118)
//   task >= E : covariance block of the unordered pair q = task - E, from hr, hc, HZ on the tape
// Both reduce to weighted moment sums over the centres,
//   A1 = sum_n u_n z_n z_n',  A2 = sum_n v_n z_n z_n',  A3 = sum_n z_n HZ_n',  y1 = sum u_n z_n,  y2 = sum v_n z_n,
// evaluated role-parallel (one matrix entry per thread) on 64-row chunks staged in shared memory.
// -------------------------------------------------------------------------------------------------
#define TB_CHUNK 64
#define TB_THREADS 256

template <int DP>
__global__ void __launch_bounds__(TB_THREADS, 3) mm_tape_bfinish_kernel(MMTapeBwd bp) {
    extern __shared__ __align__(16) double tb_dyn[];          // [2][np]: per-centre weights u, v
    const pilco_gp_model& gp = bp.gp;
    const MMTapeL& TL = bp.TL;
    const int n = gp.n, D = gp.D, E = gp.E, np = TL.np;
    const int r = blockIdx.y, task = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NSYM = DP * (DP + 1) / 2, NROLE = NSYM + DP * DP + 2 * DP + 2;
    constexpr int ZS = DP + 1;                                  // odd row stride of the staged chunks

    __shared__ double sW[MAXD * SLD], sCm[MAXD * SLD], sT[MAXD * SLD], sX[MAXD * SLD];
    __shared__ double sinvd[MAXD], spa[MAXD], spb[MAXD], sgv[MAXD], swgv[MAXD], sm[MAXD];
    __shared__ double sA1[DP * DP], sA2[DP * DP], sA3[DP * DP], sy1[DP], sy2[DP], ssum[2], sscal[4];
    __shared__ double sZc[TB_CHUNK * ZS], sHc[TB_CHUNK * ZS];
    double* su = tb_dyn;
    double* sv = tb_dyn + np;

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
    if (tid < DP) sm[tid] = tid < D ? mr[tid] : 0.0;

    const bool is_out = task < E;
    int a = task, b = task, q = 0;
    const double* HZg = nullptr;
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
        const double* hr = tpr + TL.hr + (size_t)q * np;
        const double* hc = tpr + TL.hc + (size_t)q * TL.cs * np;
        for (int nn = tid; nn < np; nn += blockDim.x) {
            double u = 0.0, v = 0.0;
            if (nn < n) {
                u = hr[nn];
                for (int k = 0; k < TL.cs; ++k) v += hc[(size_t)k * np + nn];
            }
            su[nn] = u; sv[nn] = v;
        }
        HZg = tpr + TL.HZ + (size_t)q * np * TL.ldh;
    }

    // ---- weighted moment sums, role-parallel ----
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int c0 = 0; c0 < n; c0 += TB_CHUNK) {
        __syncthreads();                                            // su/sv written; previous chunk consumed
        for (int e = tid; e < TB_CHUNK * DP; e += blockDim.x) {
            const int k = e / DP, d = e % DP, nn = c0 + k;
            const bool in = nn < n && d < D;
            sZc[k * ZS + d] = in ? X[(size_t)nn * D + d] - sm[d] : 0.0;
            sHc[k * ZS + d] = (in && HZg) ? HZg[(size_t)nn * TL.ldh + d] : 0.0;
        }
        __syncthreads();
        const int rows = (n - c0) < TB_CHUNK ? (n - c0) : TB_CHUNK;
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
            const int role = tid + TB_THREADS * slot;
            if (role >= NROLE) break;
            double a1 = acc[slot][0], a2 = acc[slot][1];
            if (role < NSYM) {                                      // (i <= j) of A1 and A2
                int j = 0;
                while ((j + 1) * (j + 2) / 2 <= role) ++j;
                const int i = role - j * (j + 1) / 2;
                for (int k = 0; k < rows; ++k) {
                    const double pz = sZc[k * ZS + i] * sZc[k * ZS + j];
                    a1 = fma(su[c0 + k], pz, a1); a2 = fma(sv[c0 + k], pz, a2);
                }
            } else if (role < NSYM + DP * DP) {                     // A3[i][j] = sum_n z_n[i] HZ_n[j]
                const int e = role - NSYM, i = e / DP, j = e % DP;
                if (HZg) for (int k = 0; k < rows; ++k) a1 = fma(sZc[k * ZS + i], sHc[k * ZS + j], a1);
            } else if (role < NSYM + DP * DP + DP) {                // y1[i], y2[i]
                const int i = role - NSYM - DP * DP;
                for (int k = 0; k < rows; ++k) { a1 = fma(su[c0 + k], sZc[k * ZS + i], a1); a2 = fma(sv[c0 + k], sZc[k * ZS + i], a2); }
            } else if (role == NSYM + DP * DP + DP) {               // sum u, sum v
                for (int k = 0; k < rows; ++k) { a1 += su[c0 + k]; a2 += sv[c0 + k]; }
            }
            acc[slot][0] = a1; acc[slot][1] = a2;
        }
    }
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
        const int role = tid + TB_THREADS * slot;
        if (role >= NROLE) break;
        const double a1 = acc[slot][0], a2 = acc[slot][1];
        if (role < NSYM) {
            int j = 0;
            while ((j + 1) * (j + 2) / 2 <= role) ++j;
            const int i = role - j * (j + 1) / 2;
            sA1[i * DP + j] = a1; sA1[j * DP + i] = a1; sA2[i * DP + j] = a2; sA2[j * DP + i] = a2;
        } else if (role < NSYM + DP * DP) sA3[role - NSYM] = a1;
        else if (role < NSYM + DP * DP + DP) { sy1[role - NSYM - DP * DP] = a1; sy2[role - NSYM - DP * DP] = a2; }
        else if (role == NSYM + DP * DP + DP) { ssum[0] = a1; ssum[1] = a2; }
    }
    __syncthreads();

    if (is_out) {
        // gW = -0.5 A1 + sym(gV y2');  gA = -W gW W - 0.5 glogc W;  sum gzeta = -W y1 + (sum w) W gV
        const double glogc = ssum[0];
        for (int e = tid; e < DP * DP; e += blockDim.x) {
            const int i = e / DP, j = e % DP;
            sT[i * SLD + j] = -0.5 * sA1[e] + 0.5 * (sgv[i] * sy2[j] + sgv[j] * sy2[i]);
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
        if (tid < DP) {
            double v = 0.0;
            for (int j = 0; j < DP; ++j) v = fma(sW[tid * SLD + j], sy1[j], v);
            Tm[tid] = v - ssum[1] * swgv[tid];                      // -(sum_n gzeta_n)
        }
        return;
    }

    // pair task
    const double g = (a == b) ? gS[a * E + a] : gS[a * E + b] + gS[b * E + a];
    const double glogR = -0.5 * g * ssum[0];
    for (int e = tid; e < DP * DP; e += blockDim.x) {               // sT = gQ / (delta_i delta_j)
        const int i = e / DP, j = e % DP;
        double v = 0.0;
        if (i < D && j < D) {
            const double gq = g * (spa[i] * spa[j] * sA1[e] + spb[i] * spb[j] * sA2[e]
                                   + spa[i] * spb[j] * sA3[i * DP + j] + spa[j] * spb[i] * sA3[j * DP + i]);
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
    if (tid < DP) {                                                 // -(sum gzeta) = g (u - 2 delta o Q u),  u = p_a y1 + p_b y2
        double qu = 0.0;
        for (int j = 0; j < DP; ++j) qu = fma(sW[tid * SLD + j], spa[j] * sy1[j] + spb[j] * sy2[j], qu);
        const double u = spa[tid] * sy1[tid] + spb[tid] * sy2[tid];
        Tm[tid] = g * (u - 2.0 * (spa[tid] + spb[tid]) * qu);
    }
}

// sum the task partials -> gm, gs
__global__ void __launch_bounds__(128) mm_tape_breduce_kernel(MMTapeBwd bp) {
    const int D = bp.gp.D, E = bp.gp.E;
    const int ntask = E + bp.TL.P;
    const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const size_t stride = MAXD + (size_t)D * D;
    const double* part = bp.part + (size_t)r * ntask * stride;
    double* gm = bp.gm + (size_t)r * bp.gm_rs;
    double* gs = bp.gs + (size_t)r * bp.gs_rs;
    for (int d = tid; d < D; d += nt) {
        double v = 0.0;
        for (int tk = 0; tk < ntask; ++tk) v += part[(size_t)tk * stride + d];
        gm[d] = bp.accumulate ? gm[d] + v : v;
    }
    for (int e = tid; e < D * D; e += nt) {
        const int i = e / D, j = e % D;
        double v = 0.0;
        for (int tk = 0; tk < ntask; ++tk)
            v += 0.5 * (part[(size_t)tk * stride + MAXD + i * D + j] + part[(size_t)tk * stride + MAXD + j * D + i]);
        gs[e] = bp.accumulate ? gs[e] + v : v;
    }
}

int mm_tape_backward_launch(const MMTapeBwd& bp, cudaStream_t st) {
    const int E = bp.gp.E, R = bp.R;
    dim3 gf(E + bp.TL.P, R);
    const size_t smem = (size_t)2 * bp.TL.np * sizeof(double);
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {                                          // large n: static + dynamic shared memory > 48 KB
        const int big = 2 * TAPE_MAX_NP * (int)sizeof(double);
        if (cudaFuncSetAttribute(mm_tape_bfinish_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess ||
            cudaFuncSetAttribute(mm_tape_bfinish_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess)
            return PILCO_ERR_LAUNCH;
        configured = true;
    }
    switch (ksteps_of(bp.gp.D)) {
        case 1: mm_tape_bfinish_kernel<4><<<gf, TB_THREADS, smem, st>>>(bp); break;
        case 2: mm_tape_bfinish_kernel<8><<<gf, TB_THREADS, smem, st>>>(bp); break;
        case 3: mm_tape_bfinish_kernel<12><<<gf, TB_THREADS, smem, st>>>(bp); break;
        default: mm_tape_bfinish_kernel<16><<<gf, TB_THREADS, smem, st>>>(bp); break;
    }
    CUDA_LAUNCH_CHECK();
    mm_tape_breduce_kernel<<<R, 128, 0, st>>>(bp);
    CUDA_LAUNCH_CHECK();
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
    return mm_tape_backward_launch(bp, (cudaStream_t)stream);
}

}  // extern "C"
