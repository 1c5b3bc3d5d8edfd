"""GPU tests of the safe-PILCO extension (reference safe_pilco_extension/): the box-risk
rewards (``pilco_box_risk``), the multiplicative reward channel of the device rollout and its reverse sweep,
and the drop-in classes.  The reference has no tests or golden vectors for this extension (oracle/safe_port.py:
parity unpinned); the checks are against the transcription, gradients against torch autograd.  fp64, tolerances
written beside each assertion."""
import numpy as np
import pytest
import torch

from util import scaled_err, make_gp_problem

pytestmark = pytest.mark.gpu
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
INF = float("inf")


def _spd(rng, k, scale=0.3):
    a = rng.rand(k, k)
    return scale * a @ a.T + 0.1 * np.eye(k)


def test_box_risk_matches_oracle():
    from oracle import safe_port as sp
    from safe_pilco_extension.rewards_safe import RiskOfCollision, SingleConstraint, ObjectiveFunction
    from pilco.rewards import ExponentialReward
    rng = np.random.RandomState(0)
    Ds = 4
    for trial in range(4):
        m, s = 0.4 * rng.randn(1, Ds), _spd(rng, Ds)
        low, high = np.array([-0.4, -0.9]), np.array([0.5, 0.3])
        R1 = RiskOfCollision(Ds, low, high)
        v, var = R1.compute_reward(m, s)
        assert abs(float(v) - sp.risk_of_collision(m, s, low, high)[0]) < 1e-13 and abs(float(var[0]) - 1e-4) < 1e-18
        for kw in (dict(dim=1, low=-0.2), dict(dim=3, high=0.6), dict(dim=2, low=-0.5, high=0.7, inside=False),
                   dict(dim=0, low=-0.1, high=0.9)):
            v, _ = SingleConstraint(**kw).compute_reward(m, s)
            assert abs(float(v) - sp.single_constraint(m, s, **kw)[0]) < 1e-13
        # ObjectiveFunction (rewards_safe.py:60-73)
        er = ExponentialReward(Ds, t=0.2 * np.ones(Ds))
        obj = ObjectiveFunction(er, R1, mu=2.5)
        v, _ = obj.compute_reward(m, s)
        ref = float(np.asarray(er.compute_reward(m, s)[0]).item()) - 2.5 * sp.risk_of_collision(m, s, low, high)[0]
        assert abs(float(np.asarray(v).item()) - ref) < 1e-12
    with pytest.raises(Exception):
        SingleConstraint(0)


def test_box_risk_derivatives_match_autograd():
    from oracle import safe_port as sp
    from pilco_b200 import engine
    rng = np.random.RandomState(1)
    Ds, R = 5, 3
    m = 0.4 * rng.randn(R, Ds)
    s = np.stack([_spd(rng, Ds) for _ in range(R)])
    dims, lows, highs, sfac, inside = (0, 2), (-0.4, -INF), (0.5, 0.3), 2.0, False
    prm = np.array([2, 0.0, sfac, 0, lows[0], highs[0], 2, lows[1], highs[1]])
    risk, dm, dv = engine.box_risk(prm, m, s, grad=True)
    for r in range(R):
        mt, st = T(m[r:r + 1]).requires_grad_(), T(s[r]).requires_grad_()
        ref = sp.box_risk_torch(mt, st, dims, lows, highs, sfac, inside)
        gm, gs = torch.autograd.grad(ref, [mt, st])
        assert abs(float(risk[r]) - float(ref.detach())) < 1e-13
        assert scaled_err(dm[r].cpu().numpy(), gm[0].numpy()) < 1e-11
        assert scaled_err(dv[r].cpu().numpy(), torch.diagonal(gs).numpy()) < 1e-11


def _problem(kind, R, seed=9):
    from pilco_b200 import engine, _lib
    Ds, U, n, bf = 3, 2, 50, 12
    D = Ds + U
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, Ds, seed=seed)
    Y = 0.1 * Y
    gp = engine.gp_factorize(X, Y, ell, sf2, sn2)
    rng = np.random.RandomState(3)
    maxa = np.array([1.5, 0.7])
    if kind == "linear":
        W, b = rng.randn(R, U, Ds), rng.randn(R, U)
        spec = dict(kind=_lib.POLICY_LINEAR, Ds=Ds, U=U, squash=True, max_action=maxa, W=W, b=b)
        params = (W, b)
    else:
        Xc, Yc, lc = rng.randn(R, bf, Ds), 0.1 * rng.randn(R, bf, U), 1.0 + 0.1 * rng.randn(R, U, Ds)
        pgp = engine.gp_factorize(Xc, Yc, lc, np.ones((R, U)), 1e-4 * np.ones((R, U)), need_iK=False, mode=1)
        spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=maxa, gp=pgp)
        params = (Xc, Yc, lc)
    return dict(Ds=Ds, U=U, X=X, Y=Y, ell=ell, sf2=sf2, sn2=sn2, gp=gp, maxa=maxa, spec=spec, params=params,
                m0=X[0, :Ds], S0=0.05 * np.eye(Ds))


def _torch_safe_total(kind, ps, P, H, Wr, tr, box_add, box_mult, mu_obj, mu):
    """torch restatement: reward_add = ObjectiveFunction(exp reward, box_add, mu_obj); reward_mult = box_mult."""
    from oracle import torch_port as tp, safe_port as sp
    iK, beta = tp.calculate_factorizations(T(P["X"]), T(P["Y"]), T(P["ell"]), T(P["sf2"]), T(P["sn2"]))
    dyn = lambda m, s: tp.predict_given_factorizations(T(P["X"]), T(P["ell"]), T(P["sf2"]), m, s, iK, beta)
    maxa = T(P["maxa"])[None]
    if kind == "linear":
        act = lambda m, s: tp.linear_action(ps[0], ps[1][None], m, s, maxa)
    else:
        act = lambda m, s: tp.rbf_action(ps[0], ps[1], ps[2], m, s, maxa)
    radd = lambda m, s: tp.exponential_reward(m, s, T(Wr), T(tr)[None]) - mu_obj * sp.box_risk_torch(m, s, *box_add)
    rmult = lambda m, s: sp.box_risk_torch(m, s, *box_mult)
    prop = lambda m, s: tp.propagate(m, s, act, dyn)
    _, _, total = sp.safe_predict_torch(T(P["m0"])[None], T(P["S0"]), H, prop, radd, rmult, mu)
    return total[0, 0]


@pytest.mark.parametrize("kind,R", [("linear", 2), ("rbf", 2)])
def test_safe_rollout_value_and_gradient_match_autograd(kind, R):
    """ADD channel = exp reward - mu_obj * box risk, MULT channel = another box risk (SafePILCO.predict,
    safe_pilco.py:29-50): rollout value and policy gradient against torch autograd on the transcription."""
    from pilco_b200 import engine, _lib
    P = _problem(kind, R)
    Ds, H, mu, mu_obj = P["Ds"], 4, 3.0, 0.7
    rng = np.random.RandomState(5)
    Wr, tr = np.diag(rng.rand(Ds) + 0.5), 0.1 * rng.rand(Ds)
    box_add = ((0, 2), (0.1, -0.2), (0.9, 0.8), 2.0, True)
    box_mult = ((1,), (0.3,), (INF,), 1.0, True)
    prm = lambda b: np.array([len(b[0]), float(b[4]), b[3]] + [v for k in range(len(b[0])) for v in (b[0][k], b[1][k], b[2][k])])
    terms = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=Wr, t=tr),
             dict(kind=_lib.REWARD_BOX, coef=-mu_obj, W=prm(box_add), t=None),
             dict(kind=_lib.REWARD_BOX, coef=1.0, channel=_lib.CHANNEL_MULT, W=prm(box_mult), t=None)]
    plan = engine.RolloutPlan(P["gp"], P["spec"], terms, P["m0"], P["S0"], H, R=R, mult_mu=mu)
    _, _, reward = plan.forward()
    g = plan.backward()
    assert float(plan.step_risk.abs().max()) > 1e-3            # the MULT channel is exercised
    for r in range(R):
        ps = [T(p[r]).requires_grad_() for p in P["params"]]
        total = _torch_safe_total(kind, ps, P, H, Wr, tr, box_add, box_mult, mu_obj, mu)
        grads = torch.autograd.grad(total, ps)
        assert abs(float(reward[r]) - float(total.detach())) < 1e-9
        names = ["W", "b"] if kind == "linear" else ["X", "Y", "ell"]
        for nm, ref in zip(names, grads):
            err = scaled_err(g[nm][r].cpu().numpy(), ref.numpy())
            assert err < 1e-7, "%s grad err %g" % (nm, err)


def test_safe_pilco_class_predict_and_optimize():
    """Drop-in classes: SafePILCO.predict against the numpy transcription; optimize_policy must not lower the
    objective and ``mu.assign`` must take effect (safe_cars_run.py usage pattern)."""
    from oracle import python_port as pp, safe_port as sp
    from safe_pilco_extension.safe_pilco import SafePILCO
    from safe_pilco_extension.rewards_safe import RiskOfCollision
    from pilco.controllers import LinearController
    from pilco.rewards import LinearReward
    np.random.seed(2)
    Ds, U, H = 4, 1, 6
    X0 = np.random.rand(70, Ds + U)
    A = np.random.rand(Ds + U, Ds)
    Y0 = 0.1 * np.sin(X0).dot(A)
    ctrl = LinearController(Ds, U, max_action=1.0)
    Wl = np.array([1.0, 0.0, 0.0, 0.0])
    low, high = np.array([0.0, 0.1]), np.array([0.6, 0.9])
    pilco = SafePILCO((X0, Y0), controller=ctrl, horizon=H, reward_add=LinearReward(Ds, Wl),
                      reward_mult=RiskOfCollision(Ds, low, high), m_init=X0[0:1, :Ds], S_init=0.05 * np.eye(Ds), mu=-4.0)
    for mod in pilco.mgpr.models:
        mod.likelihood.variance.assign(1e-3)
        mod.kernel.lengthscales.assign(np.ones(Ds + U) * 2.0)
    M, S, total = pilco.predict(pilco.m_init, pilco.S_init, H)
    ell = np.stack([mod.kernel.lengthscales.numpy() for mod in pilco.mgpr.models])
    sf2 = np.stack([mod.kernel.variance.numpy() for mod in pilco.mgpr.models])
    sn2 = np.stack([mod.likelihood.variance.numpy() for mod in pilco.mgpr.models])
    iK, beta = pp.calculate_factorizations(X0, Y0, ell, sf2, sn2)
    act = lambda m, s: pp.linear_action(ctrl.W.numpy(), ctrl.b.numpy(), m, s, True, np.ones((1, U)))
    dyn = lambda m, s: pp.predict_given_factorizations(X0, ell, sf2, m, s, iK, beta)
    prop = lambda m, s: pp.propagate(m, s, act, dyn)
    radd = lambda m, s: (m @ Wl.reshape(Ds, 1), None)
    rmult = lambda m, s: sp.risk_of_collision(m, s, low, high)
    Mr, Sr, tot_ref = sp.safe_predict(pilco.m_init, pilco.S_init, H, prop, radd, rmult, -4.0)
    assert scaled_err(M, Mr) < 1e-9 and scaled_err(S, Sr) < 1e-8
    assert abs(float(np.asarray(total).item()) - float(tot_ref.item())) < 1e-9
    r0 = float(np.asarray(pilco.compute_reward()).item())
    pilco.optimize_policy(maxiter=15, restarts=2)
    r1 = float(np.asarray(pilco.compute_reward()).item())
    assert np.isfinite(r1) and r1 >= r0 - 1e-9
    pilco.mu.assign(0.0)
    r2 = float(np.asarray(pilco.compute_reward()).item())
    M2, S2, add_only = pilco.predict(pilco.m_init, pilco.S_init, H)
    assert abs(r2 - float(np.asarray(add_only).item())) < 1e-12 and abs(r2 - r1) > 1e-6
