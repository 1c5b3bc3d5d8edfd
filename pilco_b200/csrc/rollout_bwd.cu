// rollout_bwd.cu -- reverse sweep of the H-step cascade: d(sum_t E[r(x_t)]) / d(policy parameters).
// The reference obtains this gradient from TensorFlow autodiff through tf.while_loop
// (pilco/models/pilco.py:47-50, 84-90); here every stage has a hand-derived VJP (oracle/staged.py) and the
// forward's per-step joint Gaussians / policy moments saved by pilco_rollout_forward are re-used.
//
// The sweep is a strictly serial chain over the H steps, so its cost at small batches is (kernels per step) x
// (latency of each): the chain is kept short.
//   once      rb_reward  grid (R,H): d reward_t / d(m_t, S_t) of EVERY step -- it depends on the saved trajectory only
//   per step t = H-1 .. 0, with a tape (pilco_rollout.tape):
//     rb_dyn_finish  grid (E+P, R): tape-driven VJP of the dynamics moment match (mm_tape_bfinish_task; the glue seeds
//                    gM = gm, gS, gV = s1'(gS+gS') are formed inline); the LAST CTA of a restart to arrive then runs
//                    rb_post for it: sum of the task partials, joint / squash / linear-policy / reward VJPs ->
//                    state cotangent (gm_t, gS_t) and the seeds of the policy moment match
//     mm_backward (RBF policy, accumulating): fused setup -> btile -> bfinish(+reduce)        = 4 launches per step
//   without a tape (memory-lean): rb_pre -> mm_backward (dynamics, recomputing) -> rb_post -> mm_backward (policy)
// then, for the RBF policy, the VJP through beta = (K + sn2 I)^-1 Y.
#include "mm_backward.cuh"
#include "mm_tape.cuh"
#include "rollout.cuh"
#include "small_kernels.cuh"

extern "C" size_t pilco_mm_bwd_workspace_bytes(int n, int D, int E, int R, int need_param);
void chol_solve_vec_launch(cudaStream_t st, int batch, int n, const double* L, int ld, long long ms, int E,
                           const double* Y, long long Y_bs, long long y_es, int yinc, double* x, long long xs);

struct RbWs {            // backward workspace (offsets in doubles)
    size_t gm, gS, gMd, gSd, gVd, gmj, gsj, gsjx, gMp, gSp, gVp, gbeta, gy, rgm, rgS, cnt, dynb, polb, total;
};

static RbWs rb_ws_layout(const pilco_rollout* ro) {
    const size_t R = ro->R, H = ro->H, Ds = ro->pol.Ds, U = ro->pol.U, D = Ds + U;
    RbWs L; size_t o = 0;
    auto take = [&](size_t len) { size_t at = o; o += (R * len + 1) & ~(size_t)1; return at; };
    L.gm = take(Ds); L.gS = take(Ds * Ds);
    L.gMd = take(Ds); L.gSd = take(Ds * Ds); L.gVd = take(D * Ds);
    L.gmj = take(D); L.gsj = take(D * D); L.gsjx = take(D * D);
    L.gMp = take(U); L.gSp = take(U * U); L.gVp = take(Ds * U);
    const size_t bf = ro->pol.kind == PILCO_POLICY_RBF ? ro->pol.rbf.n : 0;
    L.gbeta = take(U * bf); L.gy = take(U * bf);
    // reward gradients of every step, per channel (0: additive, 1: multiplicative): [H][2][R][Ds], [H][2][R][Ds*Ds]
    L.rgm = take(H * 2 * Ds); L.rgS = take(H * 2 * Ds * Ds);
    L.cnt = take(1);                                           // arrival counters of rb_dyn_finish (one per restart)
    // dynamics GP: task partials of the tape-driven reverse sweep, or the workspace of the recomputing one
    L.dynb = o;
    if (ro->tape) o += ((size_t)ro->R * mm_tape_bwd_part_doubles(ro->dyn.D, ro->dyn.E) + 1) & ~(size_t)1;
    else o += pilco_mm_bwd_workspace_bytes(ro->dyn.n, ro->dyn.D, ro->dyn.E, ro->R, 0) / 8;
    L.polb = o;
    if (ro->pol.kind == PILCO_POLICY_RBF)
        o += pilco_mm_bwd_workspace_bytes(ro->pol.rbf.n, ro->pol.rbf.D, ro->pol.rbf.E, ro->R, 1) / 8;
    L.total = o;
    return L;
}

struct RbDev {
    int R, H, Ds, U, t;
    int pol_kind, squash;
    const double* maxa;
    const double* W; long long W_bs;
    int n_rewards; pilco_reward_term rewards[8];
    const double* traj_m; const double* traj_S;
    const double* risk; double mult_mu;       // MULT channel: per-step risks risk[t*R + r] saved by the forward
    // forward slots of step t
    const double *sj, *Vd, *Mp, *Sp, *Vp, *Mu, *Su, *Cq, *Vu;
    // cotangent buffers [R, len]
    double *gm, *gS, *gMd, *gSd, *gVd, *gmj, *gsj, *gsjx, *gMp, *gSp, *gVp;
    double *gW, *gb;      // linear policy gradient accumulators [R,U,Ds], [R,U]
    double *rgm, *rgS;    // reward gradients of every step (rb_reward_kernel)
    // taped dynamics VJP: task partials [R][ntask][MAXD + D*D] that rb_post sums into gmj / gsj itself (else NULL);
    // then the glue cotangent gsjx is formed in shared memory too (no rb_pre launch)
    const double* tpart; int tp_ntask;
    unsigned* cnt;
};

// shared-memory scratch of rb_post_body (bytes), carved from a dynamic buffer
struct RbPostSmem { SmallScratch sc; double gMu[MAXD], gSu[MAXD * MAXD], gVu[MAXD * MAXD], gB[MAXD * MAXD], gCd[MAXD], gsjx[MAXD * MAXD]; };

// d (sum_k coef_k reward_k)(x_t) / d(m_t, S_t) for every step t and restart r, per accumulation channel
__global__ void __launch_bounds__(128) rb_reward_kernel(RbDev p) {
    PDL_ENTRY();
    __shared__ SmallScratch sc;
    const int r = blockIdx.x, t = blockIdx.y, Ds = p.Ds;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* mx = p.traj_m + ((size_t)r * (p.H + 1) + t) * Ds;
    const double* sx = p.traj_S + ((size_t)r * (p.H + 1) + t) * Ds * Ds;
    double* gm0 = p.rgm + (((size_t)t * 2) * p.R + r) * Ds;
    double* gS0 = p.rgS + (((size_t)t * 2) * p.R + r) * Ds * Ds;
    double* gm1 = gm0 + (size_t)p.R * Ds;
    double* gS1 = gS0 + (size_t)p.R * Ds * Ds;
    for (int i = tid; i < Ds; i += nt) { gm0[i] = 0.0; gm1[i] = 0.0; }
    for (int e = tid; e < Ds * Ds; e += nt) { gS0[e] = 0.0; gS1[e] = 0.0; }
    __syncthreads();
    for (int k = 0; k < p.n_rewards; ++k) {
        const pilco_reward_term& rt = p.rewards[k];
        const bool mult = rt.channel == PILCO_CHANNEL_MULT;
        dev_reward_bwd(Ds, rt, mx, sx, rt.coef, mult ? gm1 : gm0, mult ? gS1 : gS0, sc);
    }
}

__global__ void __launch_bounds__(128) rb_pre_kernel(RbDev p) {
    PDL_ENTRY();
    const int r = blockIdx.x, Ds = p.Ds, U = p.U, D = Ds + U;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* gm = p.gm + (size_t)r * Ds;
    const double* gS = p.gS + (size_t)r * Ds * Ds;
    const double* sj = p.sj + (size_t)r * D * D;
    const double* Vd = p.Vd + (size_t)r * D * Ds;
    for (int i = tid; i < Ds; i += nt) p.gMd[(size_t)r * Ds + i] = gm[i];
    for (int e = tid; e < Ds * Ds; e += nt) p.gSd[(size_t)r * Ds * Ds + e] = gS[e];
    // gVd = s1^T (gS + gS^T),   s1 = sj[:Ds, :]
    for (int e = tid; e < D * Ds; e += nt) {
        const int k = e / Ds, j = e % Ds;
        double v = 0.0;
        for (int i = 0; i < Ds; ++i) v = fma(sj[i * D + k], gS[i * Ds + j] + gS[j * Ds + i], v);
        p.gVd[(size_t)r * D * Ds + e] = v;
    }
    // extra cotangent of the joint covariance: rows < Ds get (gS + gS^T) Vd^T
    for (int e = tid; e < D * D; e += nt) {
        const int i = e / D, k = e % D;
        double v = 0.0;
        if (i < Ds) for (int j = 0; j < Ds; ++j) v = fma(gS[i * Ds + j] + gS[j * Ds + i], Vd[k * Ds + j], v);
        p.gsjx[(size_t)r * D * D + e] = v;
    }
}

// joint / squash / linear-policy / reward VJPs of step t for restart r (all 128 threads of one CTA)
__device__ __forceinline__ void rb_post_body(const RbDev& p, int r, RbPostSmem& sm) {
    SmallScratch& sc = sm.sc;
    double* gMu = sm.gMu; double* gSu = sm.gSu; double* gVu = sm.gVu; double* gB = sm.gB; double* gCd = sm.gCd;
    const int Ds = p.Ds, U = p.U, D = Ds + U, t = p.t;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* gm = p.gm + (size_t)r * Ds;
    double* gS = p.gS + (size_t)r * Ds * Ds;
    const double* gsjx = p.gsjx + (size_t)r * D * D;
    if (p.tpart) {
        // taped dynamics VJP: sum its task partials (mm_tape_breduce folded in) and form the glue cotangent
        // gsjx = [(gS + gS') Vd' ; 0] (rb_pre folded in) -- both read gS BEFORE this step's update below
        mm_tape_reduce_device(p.tpart + (size_t)r * p.tp_ntask * (MAXD + (size_t)D * D), p.tp_ntask, D,
                              p.gmj + (size_t)r * D, p.gsj + (size_t)r * D * D, 0);
        const double* Vd = p.Vd + (size_t)r * D * Ds;
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, k = e % D;
            double v = 0.0;
            if (i < Ds) for (int j = 0; j < Ds; ++j) v = fma(gS[i * Ds + j] + gS[j * Ds + i], Vd[k * Ds + j], v);
            sm.gsjx[e] = v;
        }
        gsjx = sm.gsjx;
        __syncthreads();
    }
    const double* gmj = p.gmj + (size_t)r * D;
    const double* gsj = p.gsj + (size_t)r * D * D;
    const double* mx = p.traj_m + ((size_t)r * (p.H + 1) + t) * Ds;
    const double* sx = p.traj_S + ((size_t)r * (p.H + 1) + t) * Ds * Ds;
    const double* Vu = p.Vu + (size_t)r * Ds * U;
    const double* Vp = p.Vp + (size_t)r * Ds * U;
    const double* Cq = p.Cq + (size_t)r * U * U;
    // ---- joint VJP (pilco.py:141-144) ----
    for (int e = tid; e < Ds * U; e += nt) {                 // gB = G12 + G21^T
        const int i = e / U, j = e % U;
        gB[e] = gsj[i * D + Ds + j] + gsjx[i * D + Ds + j] + gsj[(Ds + j) * D + i] + gsjx[(Ds + j) * D + i];
    }
    for (int i = tid; i < U; i += nt) gMu[i] = gmj[Ds + i];
    for (int e = tid; e < U * U; e += nt) {
        const int i = e / U, j = e % U;
        gSu[e] = gsj[(Ds + i) * D + Ds + j] + gsjx[(Ds + i) * D + Ds + j];
    }
    __syncthreads();
    for (int i = tid; i < Ds; i += nt) gm[i] += gmj[i];
    for (int e = tid; e < Ds * Ds; e += nt) {
        const int i = e / Ds, j = e % Ds;
        double v = gsj[i * D + j] + gsjx[i * D + j];
        for (int k = 0; k < U; ++k) v = fma(gB[i * U + k], Vu[j * U + k], v);       // gB Vu^T
        gS[e] += v;
    }
    for (int e = tid; e < Ds * U; e += nt) {                 // gVu = s_x^T gB
        const int i = e / U, j = e % U;
        double v = 0.0;
        for (int k = 0; k < Ds; ++k) v = fma(sx[k * Ds + i], gB[k * U + j], v);
        gVu[e] = v;
    }
    __syncthreads();
    // ---- squash VJP (controllers.py:13-36, 118-120):  Vu = Vp C ----
    double* gMp = p.gMp + (size_t)r * U;
    double* gSp = p.gSp + (size_t)r * U * U;
    double* gVp = p.gVp + (size_t)r * Ds * U;
    if (p.squash) {
        for (int k = tid; k < U; k += nt) {
            double v = 0.0;
            for (int i = 0; i < Ds; ++i) v = fma(Vp[i * U + k], gVu[i * U + k], v);
            gCd[k] = v;
        }
        for (int e = tid; e < Ds * U; e += nt) gVp[e] = gVu[e] * Cq[(e % U) * U + (e % U)];
        __syncthreads();
        dev_squash_bwd(U, p.Mp + (size_t)r * U, p.Sp + (size_t)r * U * U, p.maxa,
                       p.Mu + (size_t)r * U, p.Su + (size_t)r * U * U, gMu, gSu, gCd, gMp, gSp);
    } else {
        for (int i = tid; i < U; i += nt) gMp[i] = gMu[i];
        for (int e = tid; e < U * U; e += nt) gSp[e] = gSu[e];
        for (int e = tid; e < Ds * U; e += nt) gVp[e] = gVu[e];
        __syncthreads();
    }
    // ---- linear policy VJP (controllers.py:52-54) ----
    if (p.pol_kind == PILCO_POLICY_LINEAR) {
        dev_linear_bwd(Ds, U, p.W + (size_t)r * p.W_bs, mx, sx, gMp, gSp, gVp,
                       p.gW + (size_t)r * U * Ds, p.gb + (size_t)r * U, gm, gS, sc);
    }
    // ---- reward VJP at state t (pilco.py:133): precomputed per channel by rb_reward_kernel ----
    // MULT channel (safe_pilco.py:44-49): d[mu (1 - prod_t' (1 - risk_t'))] / d risk_t = mu prod_{t' != t} (1 - risk_t')
    double wmult = p.mult_mu;
    if (wmult != 0.0)
        for (int tt = 0; tt < p.H; ++tt) if (tt != t) wmult *= 1.0 - p.risk[(size_t)tt * p.R + r];
    const double* ra_m = p.rgm + (((size_t)t * 2) * p.R + r) * Ds;
    const double* ra_S = p.rgS + (((size_t)t * 2) * p.R + r) * Ds * Ds;
    const double* rm_m = ra_m + (size_t)p.R * Ds;
    const double* rm_S = ra_S + (size_t)p.R * Ds * Ds;
    for (int i = tid; i < Ds; i += nt) gm[i] += ra_m[i] + wmult * rm_m[i];
    for (int e = tid; e < Ds * Ds; e += nt) gS[e] += ra_S[e] + wmult * rm_S[e];
}

__global__ void __launch_bounds__(128) rb_post_kernel(RbDev p) {
    PDL_ENTRY();
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];
    rb_post_body(p, blockIdx.x, *reinterpret_cast<RbPostSmem*>(rb_dyn_smem));
}

// taped path: dynamics VJP from the tape + (last CTA of the restart) rb_post, ONE launch per step
template <int DP>
__global__ void __launch_bounds__(TB_THREADS, 3) rb_dyn_finish_kernel(MMTapeBwd bp, RbDev d) {
    PDL_ENTRY();
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];
    const int r = blockIdx.y;
    mm_tape_bfinish_task<DP>(bp, r, blockIdx.x, reinterpret_cast<double*>(rb_dyn_smem));
    if (last_cta_arrives(d.cnt + r, gridDim.x)) rb_post_body(d, r, *reinterpret_cast<RbPostSmem*>(rb_dyn_smem));
}

template <int DP>
static int launch_dyn_finish(const MMTapeBwd& tb, const RbDev& d, cudaStream_t st) {
    size_t smem = mm_tape_bfinish_smem_bytes(tb.TL.np, tb.gp.D);
    if (smem < sizeof(RbPostSmem)) smem = sizeof(RbPostSmem);
    static bool configured_dev[PILCO_MAX_DEVICES] = {false};
    bool& configured = configured_dev[pilco_current_device()];
    if (!configured) {
        const int big = 196 * 1024;
        if (cudaFuncSetAttribute(rb_dyn_finish_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return PILCO_ERR_LAUNCH;
        configured = true;
    }
    launch_hi(rb_dyn_finish_kernel<DP>, dim3(tb.gp.E + tb.TL.P, tb.R), dim3(TB_THREADS), smem, st, tb, d);
    return PILCO_OK;
}

// VJP through beta_a = (K_a + sn2 I)^-1 y_a for the RBF policy (sf2 frozen, controllers.py:91-93): given gy_a = (K_a+sn2 I)^-1 gbeta_a,
//   gY[:,a] = gy_a ;  gK = -gy beta^T ;  gX += ..., gell += ...   (oracle/staged.py: rbf_factor_backward)
__global__ void __launch_bounds__(128) rbf_factor_bwd_kernel(int bf, int Ds, int U, const double* X, const double* ell,
                                                             const double* sf2, long long sf2_bs,
                                                             const double* beta, const double* gy,
                                                             double* gX, double* gY, double* gell) {
    PDL_ENTRY();
    // one CTA per restart, thread n owns centre n: K[n][m] is evaluated once per (n, m, a)
    const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const double* Xr = X + (size_t)r * bf * Ds;
    const double* lr = ell + (size_t)r * U * Ds;
    const double* br = beta + (size_t)r * U * bf;
    const double* gr = gy + (size_t)r * U * bf;
    __shared__ double sgl[MAXD * 4], sglo[MAXD];
    for (int e = tid; e < bf * U; e += nt) { const int n = e / U, a = e % U; gY[(size_t)r * bf * U + e] = gr[a * bf + n]; }
    for (int a = 0; a < U; ++a) {
        double il2[MAXD], accl[MAXD];
        const double sfa = sf2[(size_t)r * sf2_bs + a];
#pragma unroll
        for (int d = 0; d < MAXD; ++d) { il2[d] = d < Ds ? 1.0 / (lr[a * Ds + d] * lr[a * Ds + d]) : 0.0; accl[d] = 0.0; }
        for (int n = tid; n < bf; n += nt) {
            double xn[MAXD], accx[MAXD];
#pragma unroll
            for (int d = 0; d < MAXD; ++d) { xn[d] = d < Ds ? Xr[n * Ds + d] : 0.0; accx[d] = 0.0; }
            const double gyn = gr[a * bf + n], bn = br[a * bf + n];
            for (int m = 0; m < bf; ++m) {
                double diff[MAXD], d2 = 0.0;
#pragma unroll
                for (int d = 0; d < MAXD; ++d) {
                    diff[d] = d < Ds ? xn[d] - Xr[m * Ds + d] : 0.0;
                    d2 = fma(diff[d] * diff[d], il2[d], d2);
                }
                const double K = sfa * exp(-0.5 * d2);
                const double gK = -gyn * br[a * bf + m] * K;                // dL/dK[n][m] * K
                const double Psym = gK - gr[a * bf + m] * bn * K;          // (gK + gK^T)[n][m] * K
#pragma unroll
                for (int d = 0; d < MAXD; ++d) {
                    accx[d] = fma(-Psym * il2[d], diff[d], accx[d]);
                    accl[d] = fma(gK, diff[d] * diff[d], accl[d]);
                }
            }
#pragma unroll
            for (int d = 0; d < MAXD; ++d) if (d < Ds) gX[(size_t)r * bf * Ds + n * Ds + d] += accx[d];
        }
        block_sum<MAXD>(accl, MAXD, sgl, sglo);
        if (tid < Ds) { const double l = lr[a * Ds + tid]; gell[(size_t)r * U * Ds + a * Ds + tid] += sglo[tid] / (l * l * l); }
        __syncthreads();
    }
}

extern "C" {

size_t pilco_rollout_bwd_workspace_bytes(const pilco_rollout* ro) {
    if (!ro || ro->R < 1 || ro->H < 0) return 0;
    return rb_ws_layout(ro).total * sizeof(double);
}

int pilco_rollout_backward(const pilco_rollout* ro, const pilco_rollout_grad* g, pilco_stream_t stream) {
    int rc = ro_check(ro);
    if (rc) return rc;
    if (!g || !g->ws) return PILCO_ERR_NULL;
    const RoWs FL = ro_ws_layout(ro);
    const RbWs BL = rb_ws_layout(ro);
    if (g->ws_bytes < BL.total * sizeof(double)) return PILCO_ERR_WORKSPACE;
    if (((uintptr_t)g->ws) & 15) return PILCO_ERR_ALIGN;
    const int R = ro->R, H = ro->H, Ds = ro->pol.Ds, U = ro->pol.U, D = Ds + U;
    const bool rbf = ro->pol.kind == PILCO_POLICY_RBF;
    if (rbf) { if (!g->gXc || !g->gYc || !g->gell || !g->pol_L) return PILCO_ERR_NULL; }
    else { if (!g->gW || !g->gb) return PILCO_ERR_NULL; }
    cudaStream_t st = (cudaStream_t)stream;
    double* fws = (double*)ro->ws;
    double* bws = (double*)g->ws;
    const size_t RR = (size_t)R;
    auto slot = [&](size_t base, size_t len, int t) { return fws + base + (size_t)t * RR * len; };
    auto buf = [&](size_t base) { return bws + base; };

    // zero the state cotangent and the parameter-gradient accumulators (and the arrival counters)
    cudaMemsetAsync(buf(BL.gm), 0, sizeof(double) * ((BL.gMd - BL.gm)), st);
    cudaMemsetAsync(buf(BL.cnt), 0, sizeof(double) * (BL.dynb - BL.cnt), st);
    const int bf = rbf ? ro->pol.rbf.n : 0;
    if (rbf) {
        cudaMemsetAsync(g->gXc, 0, sizeof(double) * RR * bf * Ds, st);
        cudaMemsetAsync(g->gell, 0, sizeof(double) * RR * U * Ds, st);
        cudaMemsetAsync(buf(BL.gbeta), 0, sizeof(double) * RR * U * bf, st);
    } else {
        cudaMemsetAsync(g->gW, 0, sizeof(double) * RR * U * Ds, st);
        cudaMemsetAsync(g->gb, 0, sizeof(double) * RR * U, st);
    }

    // arrival counters of the finish kernels (self-cleaning afterwards): zero once per reverse sweep
    if (!ro->tape) mm_bwd_zero_counters(mm_bws_layout(ro->dyn.n, ro->dyn.D, ro->dyn.E, 0), bws + BL.dynb, R, st);
    if (rbf) mm_bwd_zero_counters(mm_bws_layout(ro->pol.rbf.n, ro->pol.rbf.D, ro->pol.rbf.E, 1), bws + BL.polb, R, st);

    RbDev d;
    d.R = R; d.H = H; d.Ds = Ds; d.U = U;
    d.pol_kind = ro->pol.kind; d.squash = ro->pol.squash; d.maxa = ro->pol.max_action;
    d.W = ro->pol.W; d.W_bs = ro->pol.W_bs;
    d.n_rewards = ro->n_rewards;
    for (int k = 0; k < 8; ++k) d.rewards[k] = ro->rewards[k];
    d.traj_m = ro->traj_m; d.traj_S = ro->traj_S;
    d.risk = fws + FL.risk; d.mult_mu = ro_count_mult(ro) > 0 ? ro->mult_mu : 0.0;
    d.gm = buf(BL.gm); d.gS = buf(BL.gS); d.gMd = buf(BL.gMd); d.gSd = buf(BL.gSd); d.gVd = buf(BL.gVd);
    d.gmj = buf(BL.gmj); d.gsj = buf(BL.gsj); d.gsjx = buf(BL.gsjx);
    d.gMp = buf(BL.gMp); d.gSp = buf(BL.gSp); d.gVp = buf(BL.gVp);
    d.gW = g->gW; d.gb = g->gb;
    d.rgm = buf(BL.rgm); d.rgS = buf(BL.rgS);
    d.tpart = ro->tape ? bws + BL.dynb : nullptr;
    d.tp_ntask = ro->dyn.E + npairs_of(ro->dyn.E);
    d.cnt = reinterpret_cast<unsigned*>(buf(BL.cnt));
    d.t = 0;
    if (H > 0) {
        d.sj = d.Vd = d.Mp = d.Sp = d.Vp = d.Mu = d.Su = d.Cq = d.Vu = nullptr;
        launch_hi(rb_reward_kernel, dim3(R, H), dim3(128), 0, st, d);     // reward gradients of every step at once
        CUDA_LAUNCH_CHECK();
    }

    for (int t = H - 1; t >= 0; --t) {
        d.t = t;
        d.sj = slot(FL.sj, (size_t)D * D, t); d.Vd = slot(FL.Vd, (size_t)D * Ds, t);
        d.Mp = slot(FL.Mp, U, t); d.Sp = slot(FL.Sp, (size_t)U * U, t); d.Vp = slot(FL.Vp, (size_t)Ds * U, t);
        d.Mu = slot(FL.Mu, U, t); d.Su = slot(FL.Su, (size_t)U * U, t); d.Cq = slot(FL.Cq, (size_t)U * U, t);
        d.Vu = slot(FL.Vu, (size_t)Ds * U, t);
        if (ro->tape) {                              // consume the tape of step t: no exponential is recomputed
            MMTapeBwd tb;
            tb.gp = ro->dyn; tb.R = R;
            tb.m = slot(FL.mj, D, t); tb.m_rs = D; tb.s = slot(FL.sj, (size_t)D * D, t); tb.s_rs = (long long)D * D;
            tb.Mfwd = slot(FL.Md, Ds, t);
            tb.gM = d.gm; tb.gS = d.gS; tb.gV = nullptr;       // seeds straight from the state cotangent; gV = s1'(gS+gS') inline
            tb.TL = mm_tape_layout(ro->dyn.n, ro->dyn.D, ro->dyn.E, R);
            tb.tape = (const double*)ro->tape + (size_t)t * RR * tb.TL.per_r;
            tb.part = bws + BL.dynb;
            tb.gm = d.gmj; tb.gm_rs = D; tb.gs = d.gsj; tb.gs_rs = (long long)D * D; tb.accumulate = 0;
            switch (ksteps_of(D)) {
                case 1: rc = launch_dyn_finish<4>(tb, d, st); break;
                case 2: rc = launch_dyn_finish<8>(tb, d, st); break;
                case 3: rc = launch_dyn_finish<12>(tb, d, st); break;
                default: rc = launch_dyn_finish<16>(tb, d, st); break;
            }
            if (rc) return rc;
            CUDA_LAUNCH_CHECK();
        } else {
            launch_hi(rb_pre_kernel, dim3(R), dim3(128), 0, st, d);
            CUDA_LAUNCH_CHECK();
            MMBwdParams bp = mm_bwd_params(&ro->dyn, R, slot(FL.mj, D, t), D, slot(FL.sj, (size_t)D * D, t), (long long)D * D,
                                           slot(FL.Md, Ds, t), d.gMd, d.gSd, d.gVd,
                                           d.gmj, D, d.gsj, (long long)D * D, nullptr, nullptr, nullptr, 0, bws + BL.dynb);
            rc = mm_backward_launch(bp, st);
            if (rc) return rc;
            launch_hi(rb_post_kernel, dim3(R), dim3(128), sizeof(RbPostSmem), st, d);
            CUDA_LAUNCH_CHECK();
        }
        if (rbf) {
            MMBwdParams pp = mm_bwd_params(&ro->pol.rbf, R, ro->traj_m + (size_t)t * Ds, (long long)(H + 1) * Ds,
                                           ro->traj_S + (size_t)t * Ds * Ds, (long long)(H + 1) * Ds * Ds,
                                           slot(FL.Mp, U, t), d.gMp, d.gSp, d.gVp,
                                           d.gm, Ds, d.gS, (long long)Ds * Ds,
                                           g->gXc, buf(BL.gbeta), g->gell, 1, bws + BL.polb);
            rc = mm_backward_launch(pp, st);
            if (rc) return rc;
        }
    }
    if (rbf) {
        // gy = (K + sn2 I)^-1 gbeta  with the Cholesky factors kept by pilco_gp_factorize
        const int ldw = pad64(bf);
        chol_solve_vec_launch(st, R * U, bf, g->pol_L, ldw, (long long)ldw * ldw, U,
                              buf(BL.gbeta), (long long)U * bf, bf, 1, buf(BL.gy), bf);
        launch_hi(rbf_factor_bwd_kernel, dim3(R), dim3(128), 0, st, bf, Ds, U, ro->pol.rbf.X, ro->pol.rbf.ell, ro->pol.rbf.sf2, ro->pol.rbf.sf2_bs,
                                                  ro->pol.rbf.beta,
                                                  buf(BL.gy), g->gXc, g->gYc, g->gell);
        CUDA_LAUNCH_CHECK();
    }
    if (g->gm0) cudaMemcpyAsync(g->gm0, d.gm, sizeof(double) * RR * Ds, cudaMemcpyDeviceToDevice, st);
    if (g->gS0) cudaMemcpyAsync(g->gS0, d.gS, sizeof(double) * RR * Ds * Ds, cudaMemcpyDeviceToDevice, st);
    return PILCO_OK;
}

}  // extern "C"
