// rollout.cu -- H-step moment-matching cascade (forward), one batch of R independent rollouts.
// Replaces PILCO.predict / PILCO.propagate (pilco/models/pilco.py:118-153; pred.m, propagate.m).
//
// Launch sequence per step t (all on one stream, graph-capturable, no host sync):
//   ro_state<t>   grid R : [t>0: finish dynamics MM of step t-1 + next-state glue] -> traj[t];
//                          linear policy + squash + joint Gaussian
//   (RBF policy)  mm_setup/mm_tile on the policy GP, then ro_policy<t>: finish + squash + joint
//   mm_setup / mm_tile on the dynamics GP with the joint Gaussian of step t
// and a final ro_state<H> that closes the last step.  The expected rewards depend on the trajectory only, so they are
// NOT on this chain (an LU factorisation per step and restart would sit on the serial path of every step): one
// ro_reward launch over all (t, r) after the last step, then ro_reward_sum adds them in step order (same summation
// order, hence the same bits, as a running sum).
#include "rollout.cuh"
#include "mm_tape.cuh"
#include "small_kernels.cuh"

struct RoDev {
    int R, H, Ds, U, t;
    int pol_kind, squash;
    const double* maxa;
    const double* W; long long W_bs; const double* b; long long b_bs;
    int n_rewards; pilco_reward_term rewards[8];
    const double* m0; long long m0_bs; const double* S0; long long S0_bs;
    double* traj_m; double* traj_S; double* reward; double* step_reward;
    double* risk; double* step_risk; double mult_mu; int n_mult;     // MULT channel (SafePILCO): risk[t*R + r]
    double* rew;                                                      // additive channel per step: rew[t*R + r]
    // per-step slots [R, len] for step t (cur) and t-1 (prev)
    double *mj, *sj, *Mp, *Sp, *Vp, *Mu, *Su, *Cq, *Vu;
    double *mj_prev, *sj_prev, *Md_prev, *Sd_prev, *Vd_prev;
    MMParams dyn_prev;        // finish of the dynamics MM launched at step t-1
    MMParams pol;             // finish of the policy MM launched at step t (RBF)
};

// squash (optional) + joint, shared by the linear and RBF paths
__device__ __forceinline__ void ro_action_tail(const RoDev& p, int r, const double* mx, const double* sx,
                                               SmallScratch& sc) {
    const int Ds = p.Ds, U = p.U, D = Ds + U;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Mp = p.Mp + (size_t)r * U;  double* Sp = p.Sp + (size_t)r * U * U;  double* Vp = p.Vp + (size_t)r * Ds * U;
    double* Mu = p.Mu + (size_t)r * U;  double* Su = p.Su + (size_t)r * U * U;
    double* Cq = p.Cq + (size_t)r * U * U;  double* Vu = p.Vu + (size_t)r * Ds * U;
    if (p.squash) {
        dev_squash_sin(U, Mp, Sp, p.maxa, Mu, Su, Cq);
        for (int e = tid; e < Ds * U; e += nt) {
            const int i = e / U, j = e % U;
            double v = 0.0;
            for (int k = 0; k < U; ++k) v = fma(Vp[i * U + k], Cq[k * U + j], v);
            Vu[e] = v;
        }
    } else {
        for (int i = tid; i < U; i += nt) Mu[i] = Mp[i];
        for (int e = tid; e < U * U; e += nt) { Su[e] = Sp[e]; Cq[e] = (e / U == e % U) ? 1.0 : 0.0; }
        for (int e = tid; e < Ds * U; e += nt) Vu[e] = Vp[e];
    }
    __syncthreads();
    dev_joint(Ds, U, mx, sx, Mu, Su, Vu, p.mj + (size_t)r * D, p.sj + (size_t)r * D * D, sc);
}

__global__ void __launch_bounds__(128) ro_state_kernel(RoDev p) {
    PDL_ENTRY();
    __shared__ SmallScratch sc;
    const int r = blockIdx.x, t = p.t, Ds = p.Ds, U = p.U, D = Ds + U;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* mx = p.traj_m + ((size_t)r * (p.H + 1) + t) * Ds;
    double* sx = p.traj_S + ((size_t)r * (p.H + 1) + t) * Ds * Ds;
    if (t == 0) {
        for (int i = tid; i < Ds; i += nt) mx[i] = p.m0[(size_t)r * p.m0_bs + i];
        for (int e = tid; e < Ds * Ds; e += nt) sx[e] = p.S0[(size_t)r * p.S0_bs + e];
        __syncthreads();
    } else {
        mm_finish_device(p.dyn_prev, r);
        __syncthreads();
        const double* mxp = mx - Ds;
        const double* sxp = sx - Ds * Ds;
        dev_glue(Ds, U, mxp, sxp, p.sj_prev + (size_t)r * D * D,
                 p.Md_prev + (size_t)r * Ds, p.Sd_prev + (size_t)r * Ds * Ds, p.Vd_prev + (size_t)r * D * Ds,
                 mx, sx, sc);
    }
    if (t >= p.H) return;
    if (p.pol_kind == PILCO_POLICY_LINEAR) {
        dev_linear_action(Ds, U, p.W + (size_t)r * p.W_bs, p.b + (size_t)r * p.b_bs, mx, sx,
                          p.Mp + (size_t)r * U, p.Sp + (size_t)r * U * U, p.Vp + (size_t)r * Ds * U, sc);
        ro_action_tail(p, r, mx, sx, sc);
    }
}

__global__ void __launch_bounds__(128) ro_policy_kernel(RoDev p) {
    PDL_ENTRY();
    __shared__ SmallScratch sc;
    const int r = blockIdx.x, t = p.t, Ds = p.Ds;
    mm_finish_device(p.pol, r);
    __syncthreads();
    const double* mx = p.traj_m + ((size_t)r * (p.H + 1) + t) * Ds;
    const double* sx = p.traj_S + ((size_t)r * (p.H + 1) + t) * Ds * Ds;
    ro_action_tail(p, r, mx, sx, sc);
}

// expected reward at the pre-step state of every step (pilco.py:130-134), CTA = (step t, restart r)
__global__ void __launch_bounds__(128) ro_reward_kernel(RoDev p) {
    PDL_ENTRY();
    __shared__ SmallScratch sc;
    const int t = blockIdx.x, r = blockIdx.y, Ds = p.Ds;
    const double* mx = p.traj_m + ((size_t)r * (p.H + 1) + t) * Ds;
    const double* sx = p.traj_S + ((size_t)r * (p.H + 1) + t) * Ds * Ds;
    double rew = 0.0, risk = 0.0;
    for (int k = 0; k < p.n_rewards; ++k) {
        const pilco_reward_term& rt = p.rewards[k];
        const double mu = dev_reward_value(Ds, rt, mx, sx, sc);
        if (rt.channel == PILCO_CHANNEL_MULT) risk = fma(rt.coef, mu, risk);
        else rew = fma(rt.coef, mu, rew);
    }
    if (threadIdx.x == 0) {
        p.rew[(size_t)t * p.R + r] = rew;
        if (p.step_reward) p.step_reward[(size_t)r * p.H + t] = rew;
        if (p.n_mult > 0) {
            p.risk[(size_t)t * p.R + r] = risk;
            if (p.step_risk) p.step_risk[(size_t)r * p.H + t] = risk;
        }
    }
}

// reward[r] = sum_t rew[t, r] in step order  (+ SafePILCO.predict, safe_pilco.py:49: mu (1 - prod_t (1 - risk_t)))
__global__ void __launch_bounds__(128) ro_reward_sum_kernel(RoDev p) {
    PDL_ENTRY();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    double s = 0.0;
    for (int tt = 0; tt < p.H; ++tt) s += p.rew[(size_t)tt * p.R + r];
    if (p.n_mult > 0) {
        double mult = 1.0;
        for (int tt = 0; tt < p.H; ++tt) mult *= 1.0 - p.risk[(size_t)tt * p.R + r];
        s += p.mult_mu * (1.0 - mult);
    }
    p.reward[r] = s;
}

int ro_check(const pilco_rollout* ro) {
    if (!ro) return PILCO_ERR_NULL;
    if (ro->R < 1 || ro->H < 0) return PILCO_ERR_DIM;
    int rc = mm_check_model(&ro->dyn);
    if (rc) return rc;
    const int Ds = ro->pol.Ds, U = ro->pol.U;
    if (Ds < 1 || U < 1 || Ds + U != ro->dyn.D || ro->dyn.E != Ds) return PILCO_ERR_DIM;
    if (ro->pol.squash && !ro->pol.max_action) return PILCO_ERR_NULL;
    if (ro->pol.kind == PILCO_POLICY_LINEAR) { if (!ro->pol.W || !ro->pol.b) return PILCO_ERR_NULL; }
    else if (ro->pol.kind == PILCO_POLICY_RBF) {
        rc = mm_check_model(&ro->pol.rbf);
        if (rc) return rc;
        if (ro->pol.rbf.D != Ds || ro->pol.rbf.E != U || ro->pol.rbf.mode != 1) return PILCO_ERR_DIM;
    } else return PILCO_ERR_UNSUPPORTED;
    if (ro->n_rewards < 1 || ro->n_rewards > 8) return PILCO_ERR_DIM;
    for (int k = 0; k < ro->n_rewards; ++k) {
        if (!ro->rewards[k].W) return PILCO_ERR_NULL;
        const int kind = ro->rewards[k].kind, ch = ro->rewards[k].channel;
        if (kind == PILCO_REWARD_EXP && !ro->rewards[k].t) return PILCO_ERR_NULL;
        if (kind != PILCO_REWARD_EXP && kind != PILCO_REWARD_LINEAR && kind != PILCO_REWARD_BOX) return PILCO_ERR_UNSUPPORTED;
        if (ch != PILCO_CHANNEL_ADD && ch != PILCO_CHANNEL_MULT) return PILCO_ERR_UNSUPPORTED;
    }
    if (!ro->m0 || !ro->S0 || !ro->traj_m || !ro->traj_S || !ro->reward || !ro->ws) return PILCO_ERR_NULL;
    return PILCO_OK;
}

extern "C" {

size_t pilco_rollout_workspace_bytes(const pilco_rollout* ro) {
    if (!ro || ro->R < 1 || ro->H < 0) return 0;
    return ro_ws_layout(ro).total * sizeof(double);
}

size_t pilco_rollout_tape_bytes(const pilco_rollout* ro) {
    if (!ro || ro->R < 1 || ro->H < 0) return 0;
    if (!mm_tape_supported(ro->dyn.n, ro->dyn.D) || ro->dyn.D < 1 || ro->dyn.D > MAXD || ro->dyn.E < 1 || ro->dyn.E > MAXE) return 0;
    return mm_tape_layout(ro->dyn.n, ro->dyn.D, ro->dyn.E, ro->R).per_r * (size_t)ro->R * (size_t)ro->H * sizeof(double);
}

int pilco_rollout_forward(const pilco_rollout* ro, pilco_stream_t stream) {
    int rc = ro_check(ro);
    if (rc) return rc;
    const RoWs L = ro_ws_layout(ro);
    if (ro->ws_bytes < L.total * sizeof(double)) return PILCO_ERR_WORKSPACE;
    if (((uintptr_t)ro->ws) & 15) return PILCO_ERR_ALIGN;
    if (ro->tape) {
        const size_t need = pilco_rollout_tape_bytes(ro);
        if (need == 0 && ro->H > 0) return PILCO_ERR_UNSUPPORTED;
        if (ro->tape_bytes < need) return PILCO_ERR_WORKSPACE;
        if (((uintptr_t)ro->tape) & 15) return PILCO_ERR_ALIGN;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = (double*)ro->ws;
    const int R = ro->R, H = ro->H, Ds = ro->pol.Ds, U = ro->pol.U, D = Ds + U;
    const size_t RR = (size_t)R;

    RoDev d;
    d.R = R; d.H = H; d.Ds = Ds; d.U = U;
    d.pol_kind = ro->pol.kind; d.squash = ro->pol.squash; d.maxa = ro->pol.max_action;
    d.W = ro->pol.W; d.W_bs = ro->pol.W_bs; d.b = ro->pol.b; d.b_bs = ro->pol.b_bs;
    d.n_rewards = ro->n_rewards;
    for (int k = 0; k < 8; ++k) d.rewards[k] = ro->rewards[k];
    d.m0 = ro->m0; d.m0_bs = ro->m0_bs; d.S0 = ro->S0; d.S0_bs = ro->S0_bs;
    d.traj_m = ro->traj_m; d.traj_S = ro->traj_S; d.reward = ro->reward; d.step_reward = ro->step_reward;
    d.risk = ws + L.risk; d.rew = ws + L.rew; d.step_risk = ro->step_risk; d.mult_mu = ro->mult_mu; d.n_mult = ro_count_mult(ro);

    auto slot = [&](size_t base, size_t len, int t) { return ws + base + (size_t)t * RR * len; };
    auto dyn_params = [&](int t) {
        MMParams p;
        p.gp = ro->dyn; p.R = R;
        p.m = slot(L.mj, D, t); p.s = slot(L.sj, (size_t)D * D, t); p.m_rs = D; p.s_rs = (long long)D * D;
        p.M = slot(L.Md, Ds, t); p.S = slot(L.Sd, (size_t)Ds * Ds, t); p.V = slot(L.Vd, (size_t)D * Ds, t);
        p.info = ro->info; p.ws = ws + L.dynws; p.L = mm_ws_layout(ro->dyn.n, ro->dyn.D, ro->dyn.E); p.bwd = 0; p.oQ = p.oC = p.oLd = 0;
        if (ro->tape) {                                  // taped tile pass: slot (t, :) of the rollout's tape
            p.TL = mm_tape_layout(ro->dyn.n, ro->dyn.D, ro->dyn.E, R);
            p.tape = (double*)ro->tape + (size_t)t * RR * p.TL.per_r;
        }
        return p;
    };
    auto pol_params = [&](int t) {
        MMParams p;
        p.gp = ro->pol.rbf; p.R = R;
        p.m = ro->traj_m + (size_t)t * Ds; p.s = ro->traj_S + (size_t)t * Ds * Ds;
        p.m_rs = (long long)(H + 1) * Ds; p.s_rs = (long long)(H + 1) * Ds * Ds;
        p.M = slot(L.Mp, U, t); p.S = slot(L.Sp, (size_t)U * U, t); p.V = slot(L.Vp, (size_t)Ds * U, t);
        p.info = ro->info; p.ws = ws + L.polws; p.L = mm_ws_layout(ro->pol.rbf.n, ro->pol.rbf.D, ro->pol.rbf.E); p.bwd = 0; p.oQ = p.oC = p.oLd = 0;
        return p;
    };

    if (ro->info) cudaMemsetAsync(ro->info, 0, sizeof(int) * RR, st);
    for (int t = 0; t <= H; ++t) {
        d.t = t;
        const int tc = t < H ? t : H - 1 < 0 ? 0 : H - 1;      // slot used for "current" pointers (unused at t==H)
        d.mj = slot(L.mj, D, tc); d.sj = slot(L.sj, (size_t)D * D, tc);
        d.Mp = slot(L.Mp, U, tc); d.Sp = slot(L.Sp, (size_t)U * U, tc); d.Vp = slot(L.Vp, (size_t)Ds * U, tc);
        d.Mu = slot(L.Mu, U, tc); d.Su = slot(L.Su, (size_t)U * U, tc); d.Cq = slot(L.Cq, (size_t)U * U, tc);
        d.Vu = slot(L.Vu, (size_t)Ds * U, tc);
        if (t > 0) {
            d.mj_prev = slot(L.mj, D, t - 1); d.sj_prev = slot(L.sj, (size_t)D * D, t - 1);
            d.Md_prev = slot(L.Md, Ds, t - 1); d.Sd_prev = slot(L.Sd, (size_t)Ds * Ds, t - 1);
            d.Vd_prev = slot(L.Vd, (size_t)D * Ds, t - 1);
            d.dyn_prev = dyn_params(t - 1);
        } else {
            d.mj_prev = d.sj_prev = d.Md_prev = d.Sd_prev = d.Vd_prev = nullptr;
            d.dyn_prev = dyn_params(0);
        }
        if (ro->pol.kind == PILCO_POLICY_RBF && t < H) d.pol = pol_params(t); else d.pol = d.dyn_prev;
        launch_hi(ro_state_kernel, dim3(R), dim3(128), 0, st, d);
        CUDA_LAUNCH_CHECK();
        if (t == H) break;
        if (ro->pol.kind == PILCO_POLICY_RBF) {
            rc = mm_forward_launch(d.pol, st, false);
            if (rc) return rc;
            launch_hi(ro_policy_kernel, dim3(R), dim3(128), 0, st, d);
            CUDA_LAUNCH_CHECK();
        }
        rc = mm_forward_launch(dyn_params(t), st, false);
        if (rc) return rc;
    }
    if (H > 0) {
        launch_hi(ro_reward_kernel, dim3(H, R), dim3(128), 0, st, d);
        CUDA_LAUNCH_CHECK();
    }
    launch_hi(ro_reward_sum_kernel, dim3((R + 127) / 128), dim3(128), 0, st, d);
    CUDA_LAUNCH_CHECK();
    return PILCO_OK;
}

}  // extern "C"
