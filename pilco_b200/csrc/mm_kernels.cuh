// mm_kernels.cuh -- device side of the fused SE-ARD moment match (forward).
//
// Algorithm (per batch element r; see DESIGN.md "Moment-match kernels"):
//   setup   grid (E + P, R): output tasks a<E   : W_a=(s+diag ell_a^2)^-1 (Cholesky), mean M_a, V_a
//                            pair tasks (a<=b)  : Q_ab = 0.5 R^-1 s via chol(s+diag(1/(p_a+p_b))),
//                                                 per-centre A'_ab[n], B_ab[n], U'_ab[n][:]
//   tile    grid (NB, P, R): T_ab = sum_{n,m} (beta_a[n] beta_b[m] - d_ab iK_a[n,m])
//                                              * exp(A'[n] + B[m] + U'[n].zeta[m])
//                            8 rows per warp, DMMA m8n8k4 for U'.zeta, table exp, TMA-staged columns
//   finish  grid (R)       : S_ab = T_ab + d_ab*diag_add - M_a M_b
// Reference: pilco/models/mgpr.py:91-149 (gp0.m:63-104, gp2.m:69-106).
#pragma once
#include "common.cuh"

struct MMWs {            // workspace layout, offsets in doubles relative to the per-restart base
    size_t zeta, betap, Ap, Bq, U, Tpart, per_r;
    int np, ldz, P, NB;
};

static inline __host__ __device__ MMWs mm_ws_layout(int n, int D, int E, bool ordered = false) {
    MMWs L;
    L.np = pad64(n); L.ldz = ldz_of(D); L.P = ordered ? E * E : npairs_of(E); L.NB = L.np / 64;
    size_t o = 0;
    L.zeta = o;  o += (size_t)L.np * L.ldz;
    L.betap = o; o += (size_t)E * L.np;
    L.Ap = o;    o += (size_t)L.P * L.np;
    L.Bq = o;    o += (size_t)L.P * L.np;
    L.U = o;     o += (size_t)L.P * L.np * L.ldz;
    L.Tpart = o; o += (size_t)L.P * L.NB;
    L.per_r = (o + 1) & ~(size_t)1;
    return L;
}

struct MMParams {
    pilco_gp_model gp;
    int R;
    const double* m; const double* s;      // [R] rows with strides m_rs / s_rs (doubles)
    long long m_rs, s_rs;
    double* M; double* S; double* V;
    int* info;
    double* ws;
    MMWs L;
    // backward-mode setup (ordered pairs q = a*E + b): also store Q_ab, C_ab=(s+diag(1/delta))^-1 and
    // logdet R_ab per pair; offsets (doubles, per restart) of those arrays inside ws.  bwd==0: unused.
    int bwd;
    size_t oQ, oC, oLd;
};

#define TILE_CM 512        // columns staged in shared memory per chunk

#ifdef __CUDACC__

// -------------------------------------------------------------------------------------------------
// setup
// -------------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ void load_sym_s(const double* __restrict__ s, int D, double* s_s) {
    // s_s[i][j] = 0.5 (s[i][j] + s[j][i]), zero padded to DP
    for (int e = threadIdx.x; e < DP * DP; e += blockDim.x) {
        const int i = e / DP, j = e % DP;
        s_s[i * SLD + j] = (i < D && j < D) ? 0.5 * (s[i * D + j] + s[j * D + i]) : 0.0;
    }
}

template <int DP, bool BWD>
__global__ void __launch_bounds__(128) mm_setup_kernel(MMParams p) {
    const int r = blockIdx.y, task = BWD ? blockIdx.x + p.gp.E : blockIdx.x;   // BWD: pair tasks only
    const pilco_gp_model& gp = p.gp;
    const int n = gp.n, D = gp.D, E = gp.E;
    const MMWs& L = p.L;
    const int np = L.np, ldz = L.ldz;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ double s_s[MAXD * SLD];      // symmetrised input covariance
    __shared__ double sA[MAXD * SLD];       // matrix being factored
    __shared__ double sB[MAXD * SLD];       // W_a or Q_ab
    __shared__ double sC[BWD ? MAXD * SLD : 1];   // BWD: (s + diag(1/delta))^-1
    __shared__ double sinvd[MAXD];
    __shared__ double sm[MAXD], spa[MAXD], spb[MAXD], sdinv[MAXD];
    __shared__ double sred[(MAXD + 1) * 4], sout[MAXD + 1];
    __shared__ double slogdet;
    __shared__ int sok;

    const double* X = gp.X + (size_t)r * gp.X_bs;
    const double* ell = gp.ell + (size_t)r * gp.ell_bs;
    const double* sf2 = gp.sf2 + (size_t)r * gp.sf2_bs;
    const double* beta = gp.beta + (size_t)r * gp.beta_bs;
    const double* mr = p.m + (size_t)r * p.m_rs;
    const double* sr = p.s + (size_t)r * p.s_rs;
    double* wsr = p.ws + (size_t)r * L.per_r;

    load_sym_s<DP>(sr, D, s_s);
    if (tid < DP) sm[tid] = tid < D ? mr[tid] : 0.0;
    __syncthreads();

    if (!BWD && task < E) {
        // ---------------- output task: mean and input-output covariance of GP a ----------------
        const int a = task;
        if (tid < DP) { const double l = tid < D ? ell[a * D + tid] : 1.0; spa[tid] = l * l; }
        __syncthreads();
        for (int e = tid; e < DP * DP; e += blockDim.x) {
            const int i = e / DP, j = e % DP;
            sA[i * SLD + j] = s_s[i * SLD + j] + (i == j ? spa[i] : 0.0);
            sB[i * SLD + j] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            const bool ok = chol_warp(sA, sinvd, DP, lane);
            chol_solve_warp(sA, sinvd, sB, DP, DP, lane);        // sB = (s + Lambda^2)^-1
            if (lane == 0) {
                double ld = chol_logdet(sinvd, DP), sl = 0.0;
                for (int d = 0; d < D; ++d) sl += log(spa[d]);
                slogdet = log(sf2[a]) + 0.5 * (sl - ld);         // log c_a   (padded dims: log 1 = 0)
                sok = ok ? 1 : 0;
            }
        }
        __syncthreads();
        if (tid == 0 && !sok && p.info) atomicOr(&p.info[r], 1);

        double acc[DP + 1];
#pragma unroll
        for (int i = 0; i <= DP; ++i) acc[i] = 0.0;
        for (int nn = tid; nn < np; nn += blockDim.x) {
            double z[DP];
            if (nn < n) {
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = d < D ? X[(size_t)nn * D + d] - sm[d] : 0.0;
                double t[DP], e = 0.0;
#pragma unroll
                for (int i = 0; i < DP; ++i) {
                    double v = 0.0;
#pragma unroll
                    for (int j = 0; j < DP; ++j) v = fma(sB[i * SLD + j], z[j], v);
                    t[i] = v; e = fma(z[i], v, e);
                }
                const double bw = beta[(size_t)a * n + nn];
                const double w = bw * exp(-0.5 * e);
                acc[DP] += w;
#pragma unroll
                for (int i = 0; i < DP; ++i) acc[i] = fma(w, t[i], acc[i]);
                wsr[L.betap + (size_t)a * np + nn] = bw;
            } else {
#pragma unroll
                for (int d = 0; d < DP; ++d) z[d] = 0.0;
                wsr[L.betap + (size_t)a * np + nn] = 0.0;
            }
            if (a == 0) {
                for (int d = 0; d < ldz; ++d) wsr[L.zeta + (size_t)nn * ldz + d] = d < DP ? z[d] : 0.0;
            }
        }
        block_sum<DP + 1>(acc, DP + 1, sred, sout);
        const double c = exp(slogdet);
        if (tid == 0) p.M[(size_t)r * E + a] = c * sout[DP];
        if (tid < D) p.V[((size_t)r * D + tid) * E + a] = c * sout[tid];
        return;
    }

    // ---------------- pair task (a <= b): Q_ab and the per-centre exponent pieces ----------------
    const int q = task - E;
    int a, b;
    if (BWD) { a = q / E; b = q % E; } else pair_decode(q, a, b);
    if (tid < DP) {
        const double la = tid < D ? ell[a * D + tid] : 1.0, lb = tid < D ? ell[b * D + tid] : 1.0;
        spa[tid] = tid < D ? 1.0 / (la * la) : 0.0;
        spb[tid] = tid < D ? 1.0 / (lb * lb) : 0.0;
        sdinv[tid] = tid < D ? 1.0 / (1.0 / (la * la) + 1.0 / (lb * lb)) : 0.0;    // 1/(p_a+p_b)
    }
    __syncthreads();
    // A = s + diag(1/(p_a+p_b));  padded dims get 1 on the diagonal (s is zero there)
    for (int e = tid; e < DP * DP; e += blockDim.x) {
        const int i = e / DP, j = e % DP;
        const double dd = (i < D) ? sdinv[i] : 1.0;
        sA[i * SLD + j] = s_s[i * SLD + j] + (i == j ? dd : 0.0);
        sB[i * SLD + j] = s_s[i * SLD + j];
        if (BWD) sC[i * SLD + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
        const bool ok = chol_warp(sA, sinvd, DP, lane);
        chol_solve_warp(sA, sinvd, sB, DP, DP, lane);            // sB = (s + Dd^-1)^-1 s
        if (BWD) chol_solve_warp(sA, sinvd, sC, DP, DP, lane);   // sC = (s + Dd^-1)^-1
        if (lane == 0) {
            double ld = chol_logdet(sinvd, DP);
            for (int d = 0; d < D; ++d) ld += log(spa[d] + spb[d]);
            slogdet = ld;                                        // log det R_ab
            sok = ok ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid == 0 && !sok && p.info) atomicOr(&p.info[r], 1);
    // Q = 0.5 R^-1 s = 0.5 diag(1/(p_a+p_b)) (s+Dd^-1)^-1 s, symmetrised into sA
    for (int e = tid; e < DP * DP; e += blockDim.x) {
        const int i = e / DP, j = e % DP;
        const double qi = sdinv[i] * sB[i * SLD + j];
        const double qj = sdinv[j] * sB[j * SLD + i];
        sA[i * SLD + j] = 0.25 * (qi + qj);
    }
    __syncthreads();
    if (BWD) {
        double* Qo = wsr + p.oQ + (size_t)q * D * D;
        double* Co = wsr + p.oC + (size_t)q * D * D;
        for (int e = tid; e < D * D; e += blockDim.x) {
            const int i = e / D, j = e % D;
            Qo[e] = sA[i * SLD + j];
            Co[e] = 0.5 * (sC[i * SLD + j] + sC[j * SLD + i]);
        }
        if (tid == 0) wsr[p.oLd + q] = slogdet;
    }
    const double lsa = log(sf2[a]), lsb = log(sf2[b]);
    const double hld = 0.5 * slogdet;
    for (int nn = tid; nn < np; nn += blockDim.x) {
        double Apv = NEG_PAD, Bqv = NEG_PAD;
        double u[DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) u[d] = 0.0;
        if (nn < n) {
            double z[DP], za[DP], zb[DP];
            double ka = lsa, kb = lsb;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                z[d] = d < D ? X[(size_t)nn * D + d] - sm[d] : 0.0;
                za[d] = spa[d] * z[d]; zb[d] = spb[d] * z[d];
                ka = fma(-0.5 * za[d], z[d], ka);
                kb = fma(-0.5 * zb[d], z[d], kb);
            }
            double qa = 0.0, qb = 0.0;
#pragma unroll
            for (int i = 0; i < DP; ++i) {
                double va = 0.0, vb = 0.0;
#pragma unroll
                for (int j = 0; j < DP; ++j) {
                    const double qij = sA[i * SLD + j];
                    va = fma(qij, za[j], va);
                    vb = fma(qij, zb[j], vb);
                }
                qa = fma(za[i], va, qa);
                qb = fma(zb[i], vb, qb);
                u[i] = 2.0 * spb[i] * va;                        // U' = p_b o (2 Q z_a)
            }
            Apv = ka + qa - hld;
            Bqv = kb + qb;
        }
        if (BWD) {
            if (b == 0) wsr[L.betap + (size_t)a * np + nn] = nn < n ? beta[(size_t)a * n + nn] : 0.0;
            if (q == 0) {
                for (int d = 0; d < ldz; ++d)
                    wsr[L.zeta + (size_t)nn * ldz + d] = (nn < n && d < D) ? X[(size_t)nn * D + d] - sm[d] : 0.0;
            }
        }
        wsr[L.Ap + (size_t)q * np + nn] = Apv;
        wsr[L.Bq + (size_t)q * np + nn] = Bqv;
        double* up = wsr + L.U + ((size_t)q * np + nn) * ldz;
#pragma unroll
        for (int d = 0; d < DP; ++d) up[d] = u[d];
        for (int d = DP; d < ldz; ++d) up[d] = 0.0;
    }
}

// -------------------------------------------------------------------------------------------------
// tile kernel
// -------------------------------------------------------------------------------------------------
static inline __host__ __device__ size_t mm_tile_smem_bytes(int np, int ldz) {
    const int cm = np < TILE_CM ? np : TILE_CM;
    return (size_t)cm * ldz * 8 + (size_t)cm * 16 + EXP_TAB * 8 + 16;
}

template <int KS>
__global__ void __launch_bounds__(256, 2) mm_tile_kernel(MMParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MMWs& L = p.L;
    const int np = L.np, ldz = L.ldz;
    const int CM = np < TILE_CM ? np : TILE_CM;
    double* sZ = reinterpret_cast<double*>(smem_raw);
    double* sBq = sZ + (size_t)CM * ldz;
    double* sBe = sBq + CM;
    double* tab = sBe + CM;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tab + EXP_TAB);

    const int r = blockIdx.z, q = blockIdx.y, rb = blockIdx.x;
    int a, b;
    pair_decode(q, a, b);
    const double* wsr = p.ws + (size_t)r * L.per_r;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = rb * 64 + warp * 8;
    const int row = row0 + g;
    const bool active = row0 < p.gp.n;                 // warp-uniform

    double ua[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ua[ks] = wsr[L.U + ((size_t)q * np + row) * ldz + 4 * ks + t];
    const double Apv = wsr[L.Ap + (size_t)q * np + row];
    const double ba = wsr[L.betap + (size_t)a * np + row];
    const bool diag = (a == b) && (p.gp.mode == 0) && (p.gp.iK != nullptr);
    const double* ikrow = diag ? p.gp.iK + ((size_t)a * p.gp.ldk + row) * p.gp.ldk : nullptr;

    exp_table_init(tab);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();

    double acc = 0.0, tr = 0.0;
    unsigned phase = 0;
    for (int c0 = 0; c0 < np; c0 += CM) {
        const int cm = (np - c0) < CM ? (np - c0) : CM;
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_expect_tx(bar, (unsigned)(cm * ldz * 8 + cm * 16));
            tma_bulk_g2s(sZ, wsr + L.zeta + (size_t)c0 * ldz, (unsigned)(cm * ldz * 8), bar);
            tma_bulk_g2s(sBq, wsr + L.Bq + (size_t)q * np + c0, (unsigned)(cm * 8), bar);
            tma_bulk_g2s(sBe, wsr + L.betap + (size_t)b * np + c0, (unsigned)(cm * 8), bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        if (active) {
            for (int cg = 0; cg < cm; cg += 32) {
                double2 ik[4];
                if (diag) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        ik[j] = *reinterpret_cast<const double2*>(ikrow + c0 + cg + 8 * j + 2 * t);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = cg + 8 * j;
                    const double2 bq = *reinterpret_cast<const double2*>(sBq + col + 2 * t);
                    double e0 = Apv + bq.x, e1 = Apv + bq.y;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double bf = sZ[(size_t)(col + g) * ldz + 4 * ks + t];
                        dmma884(e0, e1, ua[ks], bf);
                    }
                    const double l0 = exp_tab(e0, tab), l1 = exp_tab(e1, tab);
                    const double2 bb = *reinterpret_cast<const double2*>(sBe + col + 2 * t);
                    acc = fma(bb.x, l0, acc);
                    acc = fma(bb.y, l1, acc);
                    if (diag) { tr = fma(ik[j].x, l0, tr); tr = fma(ik[j].y, l1, tr); }
                }
            }
        }
        __syncthreads();
    }
    // row sums -> beta_a-weighted total
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    double v = (t == 0) ? ba * acc : 0.0;
    v = warp_sum(v);
    tr = warp_sum(tr);
    __shared__ double sred[8];
    if (lane == 0) sred[warp] = active ? (v - tr) : 0.0;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += sred[w];
        p.ws[(size_t)r * L.per_r + L.Tpart + (size_t)q * L.NB + rb] = tot;
    }
}

// -------------------------------------------------------------------------------------------------
// finish
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mm_finish_device(const MMParams& p, int r) {
    const pilco_gp_model& gp = p.gp;
    const int E = gp.E;
    const MMWs& L = p.L;
    const double* wsr = p.ws + (size_t)r * L.per_r;
    const double* sf2 = gp.sf2 + (size_t)r * gp.sf2_bs;
    for (int q = threadIdx.x; q < L.P; q += blockDim.x) {
        int a, b;
        pair_decode(q, a, b);
        double T = 0.0;
        for (int k = 0; k < L.NB; ++k) T += wsr[L.Tpart + (size_t)q * L.NB + k];
        const double Ma = p.M[(size_t)r * E + a], Mb = p.M[(size_t)r * E + b];
        double v = T - Ma * Mb;
        if (a == b) v += (gp.mode == 0) ? sf2[a] : 1e-6;
        p.S[((size_t)r * E + a) * E + b] = v;
        p.S[((size_t)r * E + b) * E + a] = v;
    }
}

static __global__ void __launch_bounds__(128) mm_finish_kernel(MMParams p) {
    mm_finish_device(p, blockIdx.x);
}

#endif  // __CUDACC__

// host-side launcher shared by mm_forward.cu and rollout.cu
int mm_forward_launch(const MMParams& p, cudaStream_t st, bool with_finish);
int mm_check_model(const pilco_gp_model* gp);
