// mm_tape.cuh -- "taped" moment match: the forward tile pass that also leaves what the reverse sweep needs.
//
// The reference differentiates MGPR.predict_given_factorizations (pilco/models/mgpr.py:91-149) with TensorFlow
// autodiff, which keeps every [E,E,N,N] intermediate of every rollout step.  The dynamics GP's inputs X are
// constants of the policy objective (pilco/models/pilco.py:80-82 freezes the model), so its VJP needs only the
// cotangents of the input moments (m, s) -- and those depend on the N x N matrix H_ab = G_ab o L'_ab of a pair only
// through its row sums, its column sums and the ONE product H_ab Z  (Z = centred inputs, N x D):
//     d T_ab / dQ       = Za' diag(hr) Za + Zb' diag(hc) Zb + Za' (H Z) Pb + (Za' (H Z) Pb)'      (Za = Z diag p_a)
//     d T_ab / dm       = -(sum_n hr_n d e/d zeta_n + sum_m hc_m d e/d zeta_m)                      (linear in hr, hc)
//     d T_ab / dlogdetR = -T_ab / 2
// These are linear in the (not yet known) cotangent of S_ab, so the FORWARD tile pass can produce them unweighted
// while it has L' in registers: per 8x8 tile two/four extra DMMA (H Z) next to the exponent's KS DMMA and the exp.
// The reverse sweep of the dynamics GP then never recomputes an exponential: it is a few small (D x N)(N x D)
// products per pair (mm_tape_bfinish_kernel).  Forward + backward cost ~1.9 tile passes instead of 1 + 4.
//
// Tape of one moment-match call, per restart (doubles): Q[P,D,D], C[P,D,D] = (s + diag 1/delta)^-1, logdetR[P],
// hr[P,np], hc[P,cs,np] (cs = row splits of a pair over CTAs), HZ[P,np,ldh].
#pragma once
#include "mm_kernels.cuh"

// row splits of one pair over CTAs: 1 once pairs x restarts fill the machine twice over
static inline __host__ __device__ int mm_tape_splits(int n, int P, int R) {
    const int NO = (n + 7) / 8;
    const long long pr = (long long)P * R;
    int cs = 1;
    if (pr < 296) cs = (int)((296 + pr - 1) / pr);
    int cap = NO / 8;
    if (cap < 1) cap = 1;
    if (cap > 4) cap = 4;
    return cs < cap ? cs : cap;
}

static inline __host__ __device__ MMTapeL mm_tape_layout(int n, int D, int E, int R) {
    MMTapeL T;
    T.np = pad64(n); T.P = npairs_of(E);
    T.ldh = 8 * ((4 * ksteps_of(D) + 7) / 8);
    T.cs = mm_tape_splits(n, T.P, R);
    size_t o = 0;
    auto take = [&](size_t len) { size_t at = o; o += (len + 1) & ~(size_t)1; return at; };
    T.Q = take((size_t)T.P * D * D);
    T.C = take((size_t)T.P * D * D);
    T.Ld = take(T.P);
    T.hr = take((size_t)T.P * T.np);
    T.hc = take((size_t)T.P * T.cs * T.np);
    T.HZ = take((size_t)T.P * T.np * T.ldh);
    T.per_r = o;
    return T;
}

#define TAPE_MAX_NP 2048        // per-warp column sums live in shared memory: 8 warps x np doubles

static inline __host__ __device__ size_t mm_tape_smem_bytes(int np, int ldz) {
    const int cm = np < TILE_CM ? np : TILE_CM;
    return (size_t)cm * ldz * 8 + (size_t)cm * 16 + EXP_TAB * 8 + 16 + (size_t)8 * np * 8;
}

#ifdef __CUDACC__

// One column sweep of NR (1 or 2) row octets of one warp over the staged columns [cbeg, cend) of a chunk.
//   exponent tile  e = B[m] + U'[n].zeta_m  by KS DMMA; its B fragment takes the centres in the order
//   pi = (0,4,1,5,2,6,3,7), so the C fragment of lane (g,t) holds columns (t, 4+t) of row g -- which IS the A
//   fragment layout of the second product (k-step 0: column t, k-step 1: column 4+t): no shuffles in between.
//   w = G o L' (without the row factor rf);  row sums hl;  column partials sum_o rf_o w (butterfly over the 8 rows);
//   HZ += w . zeta[cols]  by 2 NT DMMA.
template <int KS, bool DIAG, int NR>
__device__ __forceinline__ void tape_sweep(const double* __restrict__ sZ, const double* __restrict__ sBq,
                                           const double* __restrict__ sBe, const double* __restrict__ tab,
                                           int cbeg, int cend, int c0,
                                           const double (&ua)[2][KS], const double (&am)[2], const double (&rf)[2],
                                           const double (&bn)[2], const double* const (&ikrow)[2],
                                           double (&hl)[2], double (&hz)[2][(4 * KS + 7) / 8][2],
                                           double* __restrict__ sCsw, int lane) {
    constexpr int ldz = KS == 1 ? 4 : (KS <= 3 ? 12 : 20);
    constexpr int NT = (4 * KS + 7) / 8;
    constexpr bool ONES = (KS & 1) != 0;                       // DP = 4 KS = 4 (mod 8)
    const int g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) | ((g & 1) << 2);                  // pi(g)
    for (int cg = cbeg; cg < cend; cg += 8) {
        const double bq0 = sBq[cg + t], bq1 = sBq[cg + 4 + t];
        const double bb0 = sBe[cg + t], bb1 = sBe[cg + 4 + t];
        double bf[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bf[ks] = sZ[(size_t)(cg + pg) * ldz + 4 * ks + t];
        double zb[2][NT];
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) zb[k2][nt] = sZ[(size_t)(cg + 4 * k2 + t) * ldz + 8 * nt + g];
        // DP = 4 (mod 8): column DP of the last n-tile is free -- a column of ones there makes the second product
        // deliver the row sums as HZ[:, DP] (lanes t == 2, element 0) and saves the explicit adds
        if (ONES && g == 4) { zb[0][NT - 1] = 1.0; zb[1][NT - 1] = 1.0; }
        double cs0 = 0.0, cs1 = 0.0;
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            double e0 = bq0, e1 = bq1;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) dmma884(e0, e1, ua[j][ks], bf[ks]);
            const double l0 = exp_shifted(e0, am[j], tab), l1 = exp_shifted(e1, am[j], tab);
            double w0, w1;
            if (DIAG) {
                const double ik0 = ikrow[j][c0 + cg + t], ik1 = ikrow[j][c0 + cg + 4 + t];   // zero padded: in bounds
                w0 = fma(bn[j], bb0, -ik0) * l0;
                w1 = fma(bn[j], bb1, -ik1) * l1;
            } else {
                w0 = bb0 * l0; w1 = bb1 * l1;
            }
            if (!ONES) hl[j] += w0 + w1;
            cs0 = fma(rf[j], w0, cs0); cs1 = fma(rf[j], w1, cs1);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                dmma884(hz[j][nt][0], hz[j][nt][1], w0, zb[0][nt]);
                dmma884(hz[j][nt][0], hz[j][nt][1], w1, zb[1][nt]);
            }
        }
        // column sums of this 8-column tile over the warp's rows: butterfly over g, lanes g == 0 own the result
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 4);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 4);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8);  cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16); cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
        if (g == 0) { sCsw[c0 + cg + t] += cs0; sCsw[c0 + cg + 4 + t] += cs1; }
    }
}

// CTA = (row split, pair, restart); 8 warps; warp gw of the pair takes the row octets gw, gw + GW, ... two at a time.
template <int KS, bool DIAG>
__device__ __forceinline__ void mm_tape_body(const MMParams& p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ldz = KS == 1 ? 4 : (KS <= 3 ? 12 : 20);
    constexpr int NT = (4 * KS + 7) / 8;
    const MMWs& L = p.L;
    const MMTapeL& TL = p.TL;
    const int np = L.np, n = p.gp.n;
    const int CM = np < TILE_CM ? np : TILE_CM;
    double* sZ = reinterpret_cast<double*>(smem_raw);
    double* sBq = sZ + (size_t)CM * ldz;
    double* sBe = sBq + CM;
    double* tab = sBe + CM;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tab + EXP_TAB);
    double* sCs = reinterpret_cast<double*>(bar + 2);          // [8 warps][np] column-sum partials

    const int r = blockIdx.z, q = blockIdx.y, sidx = blockIdx.x, cs = gridDim.x;
    int a, b;
    pair_decode(q, a, b);
    const double* wsr = p.ws + (size_t)r * L.per_r;
    double* tpr = p.tape + (size_t)r * TL.per_r;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int NO = (n + 7) >> 3, NS = np >> 3;                 // live row octets / partial-sum slots of a pair
    const int ncol8 = (n + 7) & ~7;
    const bool single = ncol8 <= CM;
    double* sCsw = sCs + (size_t)warp * np;

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    for (int c = lane; c < np; c += 32) sCsw[c] = 0.0;
    auto issue_chunk = [&](int c0, bool with_table) {
        const int cm = (np - c0) < CM ? (np - c0) : CM;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        mbar_expect_tx(bar, (unsigned)(cm * ldz * 8 + cm * 16 + (with_table ? EXP_TAB * 8 : 0)));
        tma_bulk_g2s(sZ, wsr + L.zeta + (size_t)c0 * ldz, (unsigned)(cm * ldz * 8), bar);
        tma_bulk_g2s(sBq, wsr + L.Bq + (size_t)q * np + c0, (unsigned)(cm * 8), bar);
        tma_bulk_g2s(sBe, wsr + L.betap + (size_t)b * np + c0, (unsigned)(cm * 8), bar);
        if (with_table) tma_bulk_g2s(tab, g_exp_tab, (unsigned)(EXP_TAB * 8), bar);
    };
    __syncthreads();                                    // barrier initialised before the first arrive / wait
    if (tid == 0) issue_chunk(0, true);

    const int GW = cs * 8, gw = sidx * 8 + warp;
    const int npass = (NS + 2 * GW - 1) / (2 * GW);     // uniform over the CTA (barriers inside when !single)
    unsigned phase = 0;
    bool staged = true;
    for (int pass = 0; pass < npass; ++pass) {
        int oct[2];
        oct[0] = gw + (2 * pass) * GW; oct[1] = gw + (2 * pass + 1) * GW;
        const int nlive = (oct[0] < NO) + (oct[1] < NO);       // warp-uniform; oct[0] < oct[1]
        double ua[2][KS], am[2], rf[2], bn[2], hl[2], hz[2][NT][2];
        const double* ikrow[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            hl[j] = 0.0; bn[j] = 0.0; rf[j] = 0.0; am[j] = EXP_MAGIC; ikrow[j] = nullptr;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { hz[j][nt][0] = hz[j][nt][1] = 0.0; }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) ua[j][ks] = 0.0;
            if (oct[j] < NO) {
                const int row = 8 * oct[j] + g;
                const double* uf = wsr + L.Ufrag + ((size_t)q * NS + oct[j]) * (KS * 32);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) ua[j][ks] = uf[ks * 32 + lane];
                double rowfac;
                exp_row_split(wsr[L.Arow + (size_t)q * np + row], am[j], rowfac);
                const double ba = wsr[L.betap + (size_t)a * np + row];
                if (DIAG) { bn[j] = ba; rf[j] = rowfac; ikrow[j] = p.gp.iK + ((size_t)a * p.gp.ldk + row) * p.gp.ldk; }
                else rf[j] = ba * rowfac;
            }
        }
        for (int c0 = 0; c0 < ncol8; c0 += CM) {
            const int cm = (np - c0) < CM ? (np - c0) : CM;
            const int cend = (ncol8 - c0) < cm ? (ncol8 - c0) : cm;
            if (staged) {
                mbar_wait(bar, phase); phase ^= 1; staged = false;
            } else if (!single) {
                __syncthreads();
                if (tid == 0) issue_chunk(c0, false);
                mbar_wait(bar, phase); phase ^= 1;
            }
            if (nlive == 2) tape_sweep<KS, DIAG, 2>(sZ, sBq, sBe, tab, 0, cend, c0, ua, am, rf, bn, ikrow, hl, hz, sCsw, lane);
            else if (nlive == 1) tape_sweep<KS, DIAG, 1>(sZ, sBq, sBe, tab, 0, cend, c0, ua, am, rf, bn, ikrow, hl, hz, sCsw, lane);
        }
        // row-side results of this pass
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (oct[j] >= NS) continue;                         // warp-uniform
            double v = 0.0;
            if (oct[j] < NO) {
                const int row = 8 * oct[j] + g;
                constexpr bool ONES = (KS & 1) != 0;
                constexpr int tsum = ONES ? 2 : 0;              // lane of the quad that holds the row sum
                double h;
                if (ONES) h = hz[j][NT - 1][0];                 // HZ[:, DP] (ones column of the second product)
                else {
                    h = hl[j];
                    h += __shfl_xor_sync(0xffffffffu, h, 1);
                    h += __shfl_xor_sync(0xffffffffu, h, 2);
                }
                const double hrv = rf[j] * h;
                if (t == tsum) { tpr[TL.hr + (size_t)q * np + row] = hrv; v = hrv; }
                double* hzrow = tpr + TL.HZ + ((size_t)q * np + row) * TL.ldh;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2*>(hzrow + 8 * nt + 2 * t) = make_double2(rf[j] * hz[j][nt][0], rf[j] * hz[j][nt][1]);
                v = warp_sum(v);
            }
            if (lane == 0) p.ws[(size_t)r * L.per_r + L.Tpart + (size_t)q * NS + oct[j]] = v;
        }
    }
    // column sums of this CTA's rows: fixed order over the warps (deterministic)
    __syncthreads();
    double* hc = tpr + TL.hc + ((size_t)q * cs + sidx) * np;
    for (int c = tid; c < np; c += blockDim.x) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += sCs[(size_t)w * np + c];
        hc[c] = v;
    }
}

// LB = the thread count promised to the compiler (the launch uses 256): with 2 CTAs per SM it caps the registers at
// 65536 / (2 LB) -- 304 -> 96 registers (2 CTAs/SM leave 16 k registers per SM for the latency-bound glue kernels of
// the other sub-batches to run BESIDE the tile CTAs), 352 -> 80 registers (3 CTAs/SM fit).  No variant spills.
template <int KS, int LB>
__global__ void __launch_bounds__(LB, 2) mm_tape_tile_kernel(MMParams p) {
    PDL_ENTRY();
    int a, b;
    pair_decode(blockIdx.y, a, b);
    if (a == b && p.gp.mode == 0 && p.gp.iK != nullptr) mm_tape_body<KS, true>(p);
    else mm_tape_body<KS, false>(p);
}

#endif  // __CUDACC__

// taped tile pass in place of mm_tile_kernel (p.tape != nullptr); defined in mm_tape.cu
int mm_tape_tile_launch(const MMParams& p, cudaStream_t st);

// reverse sweep from the tape (need_param = 0: cotangents of the input moments only)
struct MMTapeBwd {
    pilco_gp_model gp; int R;
    const double* m; const double* s; long long m_rs, s_rs;      // forward inputs
    const double* Mfwd;                                          // forward output M [R,E]
    const double* gM; const double* gS; const double* gV;        // cotangents [R,E],[R,E,E],[R,D,E]
    const double* tape; MMTapeL TL;
    double* part;                                                // [R][ntask][MAXD + D*D] task partials
    double* gm; double* gs; long long gm_rs, gs_rs;
    int accumulate;
};
static inline __host__ __device__ size_t mm_tape_bwd_part_doubles(int D, int E) {
    return (size_t)(E + npairs_of(E)) * (MAXD + (size_t)D * D);
}
// with_reduce = false: the caller's next kernel sums the task partials itself (mm_tape_reduce_device)
int mm_tape_backward_launch(const MMTapeBwd& bp, cudaStream_t st, bool with_reduce);

#ifdef __CUDACC__
// sum of the task partials of one restart -> gm [D], gs [D,D] (symmetrised); all threads of the CTA take part.
// Fixed order over the tasks (deterministic); 4 independent partial sums per element hide the load latency.
__device__ __forceinline__ void mm_tape_reduce_device(const double* __restrict__ part, int ntask, int D,
                                                      double* __restrict__ gm, double* __restrict__ gs, int accumulate) {
    const size_t stride = MAXD + (size_t)D * D;
    for (int e = threadIdx.x; e < D + D * D; e += blockDim.x) {
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        if (e < D) {
            const double* p0 = part + e;
            int tk = 0;
            for (; tk + 3 < ntask; tk += 4) {
                v0 += p0[(size_t)tk * stride]; v1 += p0[(size_t)(tk + 1) * stride];
                v2 += p0[(size_t)(tk + 2) * stride]; v3 += p0[(size_t)(tk + 3) * stride];
            }
            for (; tk < ntask; ++tk) v0 += p0[(size_t)tk * stride];
            const double v = (v0 + v1) + (v2 + v3);
            gm[e] = accumulate ? gm[e] + v : v;
        } else {
            const int k = e - D, i = k / D, j = k % D;
            const double* p0 = part + MAXD + i * D + j;
            const double* p1 = part + MAXD + j * D + i;
            int tk = 0;
            for (; tk + 1 < ntask; tk += 2) {
                v0 += p0[(size_t)tk * stride]; v1 += p1[(size_t)tk * stride];
                v2 += p0[(size_t)(tk + 1) * stride]; v3 += p1[(size_t)(tk + 1) * stride];
            }
            for (; tk < ntask; ++tk) { v0 += p0[(size_t)tk * stride]; v1 += p1[(size_t)tk * stride]; }
            const double v = 0.5 * ((v0 + v1) + (v2 + v3));
            gs[k] = accumulate ? gs[k] + v : v;
        }
    }
}
#endif
