"""GPU tests of the hand-derived reverse pass (``-m gpu``): moment-match VJP against the numpy statement
(oracle/staged.py, itself checked against torch autograd in test_oracle.py) and the full rollout gradient
against torch autograd on the reference port.  No reference test pins gradients, so these are the
parity tests for row a12 of the scope table; fp64 tolerance 1e-7 relative to the gradient scale."""
import numpy as np
import pytest
import torch

from util import scaled_err, make_gp_problem, make_input

pytestmark = pytest.mark.gpu
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)


@pytest.mark.parametrize("n,D,E,mode,R", [(40, 3, 2, 0, 1), (70, 5, 4, 0, 2), (130, 12, 3, 0, 1), (30, 3, 2, 1, 2),
                                          (50, 10, 2, 1, 1), (65, 13, 2, 1, 1), (300, 4, 2, 0, 1),
                                          (600, 3, 2, 0, 1), (600, 5, 2, 1, 1)])      # n > 512: multi-chunk columns
def test_mm_backward_matches_staged(n, D, E, mode, R):
    from oracle import python_port as pp, staged as st
    from pilco_b200 import engine
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, E, seed=n + mode)
    if mode == 1:
        sf2, sn2 = np.ones(E), 1e-4 * np.ones(E)
    gp = engine.gp_factorize(X, Y, ell, sf2, sn2, need_iK=(mode == 0), mode=mode)
    iK, beta = pp.calculate_factorizations(X, Y, ell, sf2, sn2)
    rng = np.random.RandomState(5)
    ms = [make_input(D, seed=20 + r, scale=0.5) for r in range(R)]
    m = np.concatenate([a for a, _ in ms]); s = np.stack([b for _, b in ms])
    gM, gS, gV = rng.randn(R, E), rng.randn(R, E, E), rng.randn(R, D, E)
    M, S, V, info = engine.mm_forward(gp, m, s)
    gm, gs, gX, gb, gl = engine.mm_backward(gp, m, s, M, gM, gS, gV, need_param=(mode == 1))
    for r in range(R):
        rm, rs, rX, rb, rl = st.mm_backward_staged(X, ell, sf2, beta, iK, m[r], s[r], gM[r], gS[r], gV[r], mode)
        assert scaled_err(gm[r].cpu().numpy(), rm) < 1e-7, "gm"
        assert scaled_err(gs[r].cpu().numpy(), rs) < 1e-7, "gs"
        if mode == 1:
            assert scaled_err(gX[r].cpu().numpy(), rX) < 1e-7, "gX"
            assert scaled_err(gb[r].cpu().numpy(), rb) < 1e-7, "gbeta"
            assert scaled_err(gl[r].cpu().numpy(), rl) < 1e-7, "gell"


@pytest.mark.parametrize("n,D,E,mode,R", [(40, 3, 2, 0, 1), (70, 5, 4, 0, 2), (130, 12, 3, 0, 1), (30, 3, 2, 1, 2),
                                          (300, 4, 2, 0, 1),            # 3 pairs x 1 restart: 4 row splits per pair
                                          (65, 13, 2, 0, 1), (77, 8, 3, 0, 3),
                                          (600, 3, 2, 0, 1), (600, 5, 3, 0, 2),      # n > 512: multi-chunk columns
                                          (300, 12, 10, 0, 2),           # BASELINE metric shape
                                          (500, 10, 8, 0, 1),            # swimmer
                                          (400, 7, 6, 0, 1),             # inverted double pendulum
                                          (200, 12, 10, 0, 1)])          # SMGPR: M = 200 inducing points
def test_mm_taped_matches_staged(n, D, E, mode, R):
    """Taped forward + tape-driven reverse sweep (pilco_mm_forward_taped / pilco_mm_backward_taped): forward outputs
    equal the plain forward's, (gm, gs) equal the numpy statement of the ordered-pair VJP (oracle/staged.py, itself
    checked against torch autograd) -- including the BASELINE.json shapes."""
    from oracle import python_port as pp, staged as st
    from pilco_b200 import engine
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, E, seed=n + mode)
    if mode == 1:
        sf2, sn2 = np.ones(E), 1e-4 * np.ones(E)
    gp = engine.gp_factorize(X, Y, ell, sf2, sn2, need_iK=(mode == 0), mode=mode)
    iK, beta = pp.calculate_factorizations(X, Y, ell, sf2, sn2)
    rng = np.random.RandomState(5)
    ms = [make_input(D, seed=20 + r, scale=0.5) for r in range(R)]
    m = np.concatenate([a for a, _ in ms]); s = np.stack([b for _, b in ms])
    gM, gS, gV = rng.randn(R, E), rng.randn(R, E, E), rng.randn(R, D, E)
    M0, S0, V0, _ = engine.mm_forward(gp, m, s)
    M, S, V, info, tape = engine.mm_forward_taped(gp, m, s)
    assert int(info.max().item()) == 0
    for got, ref in ((M, M0), (S, S0), (V, V0)):
        assert scaled_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-9      # (summation order differs; S carries the trace term x |iK|)
    gm, gs = engine.mm_backward_taped(gp, m, s, M, gM, gS, gV, tape)
    for r in range(R):
        rm, rs = st.mm_backward_staged(X, ell, sf2, beta, iK, m[r], s[r], gM[r], gS[r], gV[r], mode)[:2]
        assert scaled_err(gm[r].cpu().numpy(), rm) < 1e-7, "gm"
        assert scaled_err(gs[r].cpu().numpy(), rs) < 1e-7, "gs"
    # ... and the recomputing device VJP gives the same numbers
    gm2, gs2 = engine.mm_backward(gp, m, s, M, gM, gS, gV)[:2]
    assert scaled_err(gm.cpu().numpy(), gm2.cpu().numpy()) < 1e-9
    assert scaled_err(gs.cpu().numpy(), gs2.cpu().numpy()) < 1e-9


def _torch_rollout_reward(kind, params, X, Y, ell, sf2, sn2, maxa, Wr, tr, m0, S0, H):
    from oracle import torch_port as tp
    iK, beta = tp.calculate_factorizations(T(X), T(Y), T(ell), T(sf2), T(sn2))
    dyn = lambda m, s: tp.predict_given_factorizations(T(X), T(ell), T(sf2), m, s, iK, beta)
    rew = lambda m, s: tp.exponential_reward(m, s, T(Wr), T(tr)[None])
    if kind == "linear":
        W, b = params
        act = lambda m, s: tp.linear_action(W, b[None], m, s, T(maxa)[None])
    else:
        Xc, Yc, lc = params
        act = lambda m, s: tp.rbf_action(Xc, Yc, lc, m, s, T(maxa)[None])
    _, _, total = tp.predict(T(m0)[None], T(S0), H, act, dyn, rew)
    return total[0, 0]


@pytest.mark.parametrize("taped", [False, True])
@pytest.mark.parametrize("kind,R", [("linear", 1), ("linear", 3), ("rbf", 1), ("rbf", 2)])
def test_rollout_gradient_matches_autograd(kind, R, taped):
    from pilco_b200 import engine, _lib
    Ds, U, n, H, bf = 3, 2, 50, 4, 12
    D = Ds + U
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, Ds, seed=9)
    Y = 0.1 * Y
    gp = engine.gp_factorize(X, Y, ell, sf2, sn2)
    rng = np.random.RandomState(3)
    maxa = np.array([1.5, 0.7])
    Wr, tr = np.diag(rng.rand(Ds) + 0.5), 0.1 * rng.rand(Ds)
    m0, S0 = X[0, :Ds], 0.05 * np.eye(Ds)
    rew = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=Wr, t=tr)]
    if kind == "linear":
        W, b = rng.randn(R, U, Ds), rng.randn(R, U)
        spec = dict(kind=_lib.POLICY_LINEAR, Ds=Ds, U=U, squash=True, max_action=maxa, W=W, b=b)
    else:
        Xc, Yc, lc = rng.randn(R, bf, Ds), 0.1 * rng.randn(R, bf, U), 1.0 + 0.1 * rng.randn(R, U, Ds)
        pgp = engine.gp_factorize(Xc, Yc, lc, np.ones((R, U)), 1e-4 * np.ones((R, U)), need_iK=False, mode=1)
        spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=maxa, gp=pgp)
    plan = engine.RolloutPlan(gp, spec, rew, m0, S0, H, R=R, grad=taped)
    assert (plan.tape is not None) == taped
    _, _, reward = plan.forward()
    g = plan.backward()
    for r in range(R):
        if kind == "linear":
            ps = [T(W[r]).requires_grad_(), T(b[r]).requires_grad_()]
        else:
            ps = [T(Xc[r]).requires_grad_(), T(Yc[r]).requires_grad_(), T(lc[r]).requires_grad_()]
        total = _torch_rollout_reward(kind, ps, X, Y, ell, sf2, sn2, maxa, Wr, tr, m0, S0, H)
        grads = torch.autograd.grad(total, ps)
        assert abs(float(reward[r]) - float(total.detach())) < 1e-9
        names = ["W", "b"] if kind == "linear" else ["X", "Y", "ell"]
        for nm, ref in zip(names, grads):
            err = scaled_err(g[nm][r].cpu().numpy(), ref.numpy())
            assert err < 1e-7, "%s grad err %g" % (nm, err)


@pytest.mark.parametrize("shape", ["metric", "swimmer"])
def test_rollout_gradient_at_baseline_shapes(shape):
    """Policy gradient of the H-step cascade at the BASELINE.json shapes (metric: N=300, D=12, E=10, bf=50;
    swimmer: N=500, D=10, E=8, bf=40), R=2 restarts: the taped path (what optimize_policy runs) against torch
    autograd on the reference port, and against the recomputing device path."""
    from pilco_b200 import engine, _lib
    from util import make_rollout_problem
    N, Ds, U, bf, H = (300, 10, 2, 50, 10) if shape == "metric" else (500, 8, 2, 40, 6)
    R = 2
    P = make_rollout_problem(N, Ds, U, bf, R, seed=0)
    gp = engine.gp_factorize(P["X"], P["Y"], P["ell"], P["sf2"], P["sn2"])
    rew = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=P["W"], t=P["t"])]
    plans = []
    for taped in (True, False):
        pgp = engine.gp_factorize(P["Xc"], P["Yc"], P["lc"], np.ones((R, U)), 1e-4 * np.ones((R, U)), need_iK=False, mode=1)
        spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=P["maxa"], gp=pgp)
        plan = engine.RolloutPlan(gp, spec, rew, P["m0"], P["S0"], H, R=R, grad=taped)
        plan.forward()
        plans.append((plan, {k: v.clone() for k, v in plan.backward().items()}, plan.reward.clone()))
    (pt, gt, rt), (pr, gr, rr) = plans
    assert pt.tape is not None and pr.tape is None
    assert int(pt.info.max().item()) == 0
    assert scaled_err(rt.cpu().numpy(), rr.cpu().numpy()) < 1e-11
    for k in ("X", "Y", "ell"):
        assert scaled_err(gt[k].cpu().numpy(), gr[k].cpu().numpy()) < 1e-8, k
    torch.set_num_threads(min(32, torch.get_num_threads()))
    for r in range(R):
        ps = [T(P["Xc"][r]).requires_grad_(), T(P["Yc"][r]).requires_grad_(), T(P["lc"][r]).requires_grad_()]
        total = _torch_rollout_reward("rbf", ps, P["X"], P["Y"], P["ell"], P["sf2"], P["sn2"], P["maxa"], P["W"], P["t"],
                                      P["m0"], P["S0"], H)
        grads = torch.autograd.grad(total, ps)
        assert abs(float(rt[r]) - float(total.detach())) < 1e-9
        for nm, ref in zip(("X", "Y", "ell"), grads):
            err = scaled_err(gt[nm][r].cpu().numpy(), ref.numpy())
            assert err < 1e-7, "%s grad err %g (%s)" % (nm, err, shape)


def test_recipe_cascade_with_policy_optimisation():
    """tests/test_cascade.py:17-78 in full: optimize_models(restarts) + optimize_policy(restarts=5), then the
    10-step cascade against pred.m at rtol 2e-4 (BASELINE.json asks 1e-3)."""
    from oracle import matlab_port as mp
    from pilco.models.pilco import PILCO
    from util import hyp_of
    np.random.seed(0)
    d, k, horizon = 2, 1, 10
    e = np.array([[10.0]])
    X0 = np.random.rand(100, d + k)
    A = np.random.rand(d + k, d)
    Y0 = np.sin(X0).dot(A) + 1e-3 * (np.random.rand(100, d) - 0.5)
    pilco = PILCO((X0, Y0))
    pilco.controller.max_action = e
    pilco.optimize_models(restarts=2)
    r0 = float(np.asarray(pilco.compute_reward()).item())
    pilco.optimize_policy(restarts=5)
    r1 = float(np.asarray(pilco.compute_reward()).item())
    assert r1 >= r0 - 1e-9, "policy optimisation must not decrease the expected reward (%g -> %g)" % (r0, r1)
    m = np.random.rand(1, d)
    s = np.random.rand(d, d)
    s = s.dot(s.T)
    M, S, reward = pilco.predict(m, s, horizon)
    policy = dict(p=dict(w=pilco.controller.W.numpy(), b=pilco.controller.b.numpy().T), maxU=e)
    ell = np.stack([mod.kernel.lengthscales.numpy() for mod in pilco.mgpr.models])
    sf2 = np.stack([mod.kernel.variance.numpy() for mod in pilco.mgpr.models])
    sn2 = np.stack([mod.likelihood.variance.numpy() for mod in pilco.mgpr.models])
    plant = dict(angi=np.zeros(0), poli=np.arange(d) + 1, dyni=np.arange(d) + 1, difi=np.arange(d) + 1)
    Mm, Sm = mp.pred(policy, plant, dict(hyp=hyp_of(ell, sf2, sn2), inputs=X0, targets=Y0), m.T, s, horizon)
    np.testing.assert_allclose(M[0], Mm[:, -1], rtol=2e-4)
    np.testing.assert_allclose(S, Sm[:, :, -1], rtol=2e-4)


def test_rbf_policy_optimisation_improves_reward():
    from pilco.models import PILCO
    from pilco.controllers import RbfController
    from pilco.rewards import ExponentialReward
    np.random.seed(1)
    Ds, U = 3, 1
    X0 = np.random.rand(80, Ds + U)
    A = np.random.rand(Ds + U, Ds)
    Y0 = 0.1 * np.sin(X0).dot(A)
    ctrl = RbfController(Ds, U, 10, max_action=2.0)
    pilco = PILCO((X0, Y0), controller=ctrl, horizon=8, reward=ExponentialReward(Ds, t=np.array([0.5, 0.5, 0.5])),
                  m_init=X0[0:1, :Ds], S_init=0.05 * np.eye(Ds))
    for mod in pilco.mgpr.models:
        mod.likelihood.variance.assign(1e-3)
        mod.kernel.lengthscales.assign(np.ones(Ds + U) * 2.0)
    r0 = float(np.asarray(pilco.compute_reward()).item())
    pilco.optimize_policy(maxiter=15, restarts=3)
    r1 = float(np.asarray(pilco.compute_reward()).item())
    assert np.isfinite(r1) and r1 >= r0 - 1e-9


def test_policy_optimisation_in_lockstep_groups(monkeypatch):
    """16 restarts dealt to two lock-step groups (own evaluator, stream and driver thread each, half an evaluation out
    of phase) must find the same per-restart optima as one lock-step batch: the restarts are independent problems
    (pilco.py:94-108 runs them one after the other)."""
    from pilco.models import PILCO
    from pilco.controllers import RbfController
    from pilco.rewards import ExponentialReward
    from pilco_b200 import policy_opt
    Ds, U = 3, 1
    rng = np.random.RandomState(5)
    X0 = rng.rand(60, Ds + U)
    Y0 = 0.1 * np.sin(X0).dot(rng.rand(Ds + U, Ds))
    out = {}
    for groups in (1, 2):
        monkeypatch.setenv("PILCO_OPT_GROUPS", str(groups))
        np.random.seed(11)
        ctrl = RbfController(Ds, U, 6, max_action=2.0)
        pilco = PILCO((X0, Y0), controller=ctrl, horizon=6, reward=ExponentialReward(Ds, t=np.array([0.5, 0.5, 0.5])),
                      m_init=X0[0:1, :Ds], S_init=0.05 * np.eye(Ds))
        for mod in pilco.mgpr.models:
            mod.likelihood.variance.assign(1e-3)
            mod.kernel.lengthscales.assign(np.ones(Ds + U) * 2.0)
        flat, best, rewards = policy_opt.optimize(pilco, maxiter=4, restarts=16)
        assert policy_opt.LAST_STATS["groups"] == groups
        assert policy_opt.LAST_STATS["rollout_steps"] > 0
        out[groups] = (flat, best, np.array(rewards))
    np.testing.assert_allclose(out[2][2], out[1][2], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out[2][1], out[1][1], rtol=1e-6)
    np.testing.assert_allclose(out[2][0], out[1][0], rtol=1e-5, atol=1e-7)


def test_cuda_graph_replay_matches_eager():
    """The captured H-step loop (forward + reverse sweep) must reproduce the eager results bit for bit, and keep
    doing so after the policy parameters are overwritten in place."""
    from pilco_b200 import engine, _lib
    Ds, U, n, H, bf, R = 3, 1, 40, 5, 8, 3
    D = Ds + U
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, Ds, seed=4)
    gp = engine.gp_factorize(X, 0.1 * Y, ell, sf2, sn2)
    rng = np.random.RandomState(0)
    Xc, Yc, lc = rng.randn(R, bf, Ds), 0.1 * rng.randn(R, bf, U), 1.0 + 0.1 * rng.randn(R, U, Ds)
    pgp = engine.gp_factorize(Xc, Yc, lc, np.ones((R, U)), 1e-4 * np.ones((R, U)), need_iK=False, mode=1)
    spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=np.array([2.0]), gp=pgp)
    plan = engine.RolloutPlan(gp, spec, [dict(kind=_lib.REWARD_EXP, coef=1.0, W=np.eye(Ds), t=np.zeros(Ds))],
                              X[0, :Ds], 0.1 * np.eye(Ds), H, R=R, grad=True)
    plan.forward(); g0 = {k: v.clone() for k, v in plan.backward().items()}
    r0, m0 = plan.reward.clone(), plan.traj_m.clone()
    plan.capture(backward=True)
    plan.reward.zero_()
    plan.replay()
    torch.cuda.synchronize()
    assert torch.equal(plan.reward, r0) and torch.equal(plan.traj_m, m0)
    for k in ("X", "Y", "ell"):
        assert torch.equal(plan.gbuf[k], g0[k])
    # new parameters in the same buffers -> refactorise -> replay == eager
    pgp.Y.mul_(1.5)
    engine.gp_refactorize(pgp)
    plan.replay(); torch.cuda.synchronize()
    r1 = plan.reward.clone()
    plan.forward(); torch.cuda.synchronize()
    assert torch.equal(plan.reward, r1) and not torch.equal(r1, r0)
