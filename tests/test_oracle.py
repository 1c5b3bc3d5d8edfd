"""CPU tests of the oracle itself: the two independent transcriptions (MATLAB loop form / Python
vectorised form) must agree on the reference's test recipes -- the relation the reference's own tests
assert at rtol 1e-4 -- and the staged algorithm statement must agree with both and with torch autograd."""
import numpy as np
import pytest
import torch

from oracle import matlab_port as mp, python_port as pp, torch_port as tp, staged as st
from util import relerr, scaled_err, make_gp_problem, make_input, hyp_of

T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)


def _recipe(d=3, k=2, seed=0):
    np.random.seed(seed)
    X0 = np.random.rand(100, d)
    A = np.random.rand(d, k)
    Y0 = np.sin(X0).dot(A) + 1e-3 * (np.random.rand(100, k) - 0.5)
    m = np.random.rand(1, d)
    s = np.random.rand(d, d)
    s = s.dot(s.T)
    rng = np.random.RandomState(1)
    ell, sf2, sn2 = 1 + rng.rand(k, d), 0.5 + rng.rand(k), np.array([1e-3, 2e-3])[:k]
    return X0, Y0, m, s, ell, sf2, sn2


def test_gp0_vs_mgpr_port():
    X0, Y0, m, s, ell, sf2, sn2 = _recipe()
    M1, S1, V1 = mp.gp0(dict(hyp=hyp_of(ell, sf2, sn2), inputs=X0, targets=Y0), m.T, s)
    M2, S2, V2 = pp.predict_on_noisy_inputs(X0, Y0, ell, sf2, sn2, m, s)
    assert relerr(M2, M1.T) < 1e-10 and relerr(S2, S1) < 1e-9 and relerr(V2, V1) < 1e-9


def test_gp1_vs_smgpr_port():
    X0, Y0, m, s, ell, sf2, sn2 = _recipe()
    Z = np.random.RandomState(2).rand(30, 3)
    M1, S1, V1 = mp.gp1(dict(hyp=hyp_of(ell, sf2, sn2), inputs=X0, targets=Y0, induce=Z), m.T, s)
    M2, S2, V2 = pp.sparse_predict_on_noisy_inputs(X0, Z, Y0, ell, sf2, sn2, m, s)
    assert relerr(M2, M1.T) < 1e-8 and relerr(S2, S1) < 1e-7 and relerr(V2, V1) < 1e-7    # FITC is ill-conditioned


def test_gp2_vs_rbf_port():
    X0, Y0, m, s, ell, _, _ = _recipe()
    hyp = hyp_of(ell, np.ones(2), 1e-4 * np.ones(2))
    M1, S1, V1 = mp.gp2(dict(hyp=hyp, inputs=X0, targets=Y0), m.T, s)
    M2, S2, V2 = pp.rbf_action(X0, Y0, ell, m, s, squash=False)
    assert relerr(M2, M1.T) < 1e-10 and relerr(S2, S1) < 1e-9 and relerr(V2, V1) < 1e-8


def test_gsin_conlin_reward_ports():
    _, _, m, s, _, _, _ = _recipe()
    M1, S1, C1 = mp.gSin(m.T, s, 7.0)
    M2, S2, C2 = pp.squash_sin(m, s, 7.0)
    assert relerr(M2, M1.T) < 1e-14 and relerr(S2, S1) < 1e-13 and scaled_err(C2, C1) < 1e-15
    rng = np.random.RandomState(3)
    W, b = rng.rand(2, 3), rng.rand(1, 2)
    M1, S1, V1 = mp.conlin(dict(p=dict(w=W, b=b.T)), m.T, s)
    M2, S2, V2 = pp.linear_action(W, b, m, s, squash=False)
    assert relerr(M2, M1.T) < 1e-14 and relerr(S2, S1) < 1e-13 and relerr(V2, V1) < 1e-15
    mu1, dm, dS, sR1 = mp.reward(m.T, s, np.zeros((3, 1)), np.eye(3))
    mu2, sR2 = pp.exponential_reward(m, s, np.eye(3), np.zeros((1, 3)))
    assert abs(mu1 - mu2[0, 0]) < 1e-15 and abs(sR1 - sR2[0, 0]) < 1e-15
    # analytic reward derivatives of reward.m:48-49 against finite differences
    eps = 1e-6
    for i in range(3):
        mm = m.copy(); mm[0, i] += eps
        fd = (pp.exponential_reward(mm, s, np.eye(3), np.zeros((1, 3)))[0][0, 0] - mu2[0, 0]) / eps
        assert abs(fd - dm[0, i]) < 1e-5


def test_cascade_ports_agree():
    """propagate.m/pred.m vs pilco.py:118-153 over 10 steps (test_cascade.py recipe)."""
    np.random.seed(0)
    d, k = 2, 1
    X0 = np.random.rand(100, d + k)
    A = np.random.rand(d + k, d)
    Y0 = np.sin(X0).dot(A) + 1e-3 * (np.random.rand(100, d) - 0.5)
    rng = np.random.RandomState(4)
    ell, sf2, sn2 = 1 + rng.rand(d, d + k), 0.5 + rng.rand(d), np.array([1e-3, 2e-3])
    W, b, e = rng.rand(k, d), rng.rand(1, k), np.array([[10.0]])
    m = rng.rand(1, d); s = rng.rand(d, d); s = s.dot(s.T)
    plant = dict(angi=np.zeros(0), poli=np.arange(d) + 1, dyni=np.arange(d) + 1, difi=np.arange(d) + 1)
    Mm, Sm = mp.pred(dict(p=dict(w=W, b=b.T), maxU=e), plant, dict(hyp=hyp_of(ell, sf2, sn2), inputs=X0, targets=Y0), m.T, s, 10)
    iK, beta = pp.calculate_factorizations(X0, Y0, ell, sf2, sn2)
    Mp, Sp, R = pp.predict(m, s, 10, lambda m, s: pp.linear_action(W, b, m, s, True, e),
                           lambda m, s: pp.predict_given_factorizations(X0, ell, sf2, m, s, iK, beta),
                           lambda m, s: pp.exponential_reward(m, s, np.eye(d), np.zeros((1, d))))
    assert relerr(Mp[0], Mm[:, -1]) < 1e-10 and relerr(Sp, Sm[:, :, -1]) < 1e-9


def test_torch_port_matches_numpy_port():
    X0, Y0, m, s, ell, sf2, sn2 = _recipe()
    iK, beta = pp.calculate_factorizations(X0, Y0, ell, sf2, sn2)
    iKt, bt = tp.calculate_factorizations(T(X0), T(Y0), T(ell), T(sf2), T(sn2))
    assert scaled_err(bt.numpy(), beta) < 1e-9
    M1, S1, V1 = pp.predict_given_factorizations(X0, ell, sf2, m, s, iK, beta)
    M2, S2, V2 = tp.predict_given_factorizations(T(X0), T(ell), T(sf2), T(m), T(s), T(iK), T(beta))
    assert relerr(M2.numpy(), M1) < 1e-12 and relerr(S2.numpy(), S1) < 1e-10 and relerr(V2.numpy(), V1) < 1e-12


@pytest.mark.parametrize("n,D,E,mode", [(40, 3, 2, 0), (30, 5, 4, 0), (20, 3, 2, 1), (15, 5, 1, 1)])
def test_staged_forward_and_vjp(n, D, E, mode):
    """The algorithm the kernels implement: forward vs the port, VJP vs torch autograd (rtol 1e-10)."""
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, E, seed=n)
    if mode == 1:
        sf2, sn2 = np.ones(E), 1e-4 * np.ones(E)
    iK, beta = pp.calculate_factorizations(X, Y, ell, sf2, sn2)
    m, s = make_input(D, seed=3)
    m = m[0]
    M, S, V = st.mm_forward_staged(X, ell, sf2, beta, iK, m, s, mode)
    Mr, Sr, Vr = pp.predict_given_factorizations(X, ell, sf2, m[None], s, iK * (0 if mode else 1), beta)
    if mode == 1:
        Sr = Sr - np.diag(sf2 - 1e-6)
    assert scaled_err(M, Mr[0]) < 1e-12 and scaled_err(S, Sr) < 1e-10 and scaled_err(V, Vr) < 1e-12
    rng = np.random.RandomState(1)
    gM, gS, gV = rng.randn(E), rng.randn(E, E), rng.randn(D, E)
    tm, ts, tX, tb, tl = [T(a).requires_grad_() for a in (m, s, X, beta, ell)]
    Mt, St, Vt = tp.predict_given_factorizations(tX, tl, T(sf2), tm[None], 0.5 * (ts + ts.T), None if mode else T(iK), tb)
    loss = (Mt[0] * T(gM)).sum() + (St * T(gS)).sum() + (Vt * T(gV)).sum()
    g = torch.autograd.grad(loss, [tm, ts, tX, tb, tl])
    out = st.mm_backward_staged(X, ell, sf2, beta, iK, m, s, gM, gS, gV, mode)
    for mine, ref in zip(out, g):
        assert scaled_err(mine, ref.numpy()) < 1e-10


def test_staged_closed_form_vjps():
    rng = np.random.RandomState(2)
    U = 3
    m, s, e = rng.randn(U), rng.rand(U, U), np.array([1.5, 0.7, 2.0])
    s = s @ s.T
    gM, gS, gC = rng.randn(U), rng.randn(U, U), np.diag(rng.randn(U))
    tm, ts = T(m).requires_grad_(), T(s).requires_grad_()
    Mt, St, Ct = tp.squash_sin(tm[None], ts, T(e)[None])
    g = torch.autograd.grad((Mt[0] * T(gM)).sum() + (St * T(gS)).sum() + (Ct * T(gC)).sum(), [tm, ts])
    gm, gs = st.squash_backward(m, s, e, gM, gS, gC)
    assert scaled_err(gm, g[0].numpy()) < 1e-12 and scaled_err(gs + gs.T, (g[1] + g[1].T).numpy()) < 1e-12
    D = 4
    m, s, W, t = rng.randn(D), rng.rand(D, D), np.diag(rng.rand(D) + 0.5), rng.randn(D)
    s = s @ s.T
    tm, ts = T(m).requires_grad_(), T(s).requires_grad_()
    mu = tp.exponential_reward(tm[None], 0.5 * (ts + ts.T), T(W), T(t)[None])[0, 0]
    g = torch.autograd.grad(mu, [tm, ts])
    mu2, dm, dS = st.exp_reward_grad(m, s, W, t)
    assert abs(mu2 - mu.item()) < 1e-14 and scaled_err(dm, g[0].numpy()) < 1e-12 and scaled_err(dS, g[1].numpy()) < 1e-12
    bf, Ds, U = 12, 3, 2
    Xc, Yc, ell, gb = rng.randn(bf, Ds), rng.randn(bf, U), 1 + 0.1 * rng.randn(U, Ds), rng.randn(U, bf)
    tX, tY, tl = T(Xc).requires_grad_(), T(Yc).requires_grad_(), T(ell).requires_grad_()
    _, bt = tp.calculate_factorizations(tX, tY, tl, torch.ones(U, dtype=torch.float64), 1e-4 * torch.ones(U, dtype=torch.float64))
    g = torch.autograd.grad((bt * T(gb)).sum(), [tX, tY, tl])
    gX, gY, gl = st.rbf_factor_backward(Xc, Yc, ell, gb)
    assert scaled_err(gX, g[0].numpy()) < 1e-10 and scaled_err(gY, g[1].numpy()) < 1e-10 and scaled_err(gl, g[2].numpy()) < 1e-10


def test_safe_extension_ports_agree():
    """oracle/safe_port.py: numpy transcription of safe_pilco_extension (SafePILCO.predict with a RiskOfCollision
    multiplicative reward and an ObjectiveFunction additive reward) against its torch twin; the torch twin's
    gradient against finite differences.  (No reference test exists for the extension: parity unpinned.)"""
    import torch
    from oracle import python_port as pp, torch_port as tp, safe_port as sp
    from util import make_gp_problem
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    Ds, U, n, H, mu, mu_obj = 3, 1, 40, 4, 3.0, 0.7
    X, Y, ell, sf2, sn2 = make_gp_problem(n, Ds + U, Ds, seed=2)
    Y = 0.1 * Y
    rng = np.random.RandomState(0)
    W, b = rng.randn(U, Ds), rng.randn(1, U)
    maxa = np.array([[1.3]])
    Wr, tr = np.eye(Ds), 0.1 * np.ones((1, Ds))
    low, high = np.array([0.1, -0.2]), np.array([0.9, 0.8])
    m0, S0 = X[0:1, :Ds], 0.05 * np.eye(Ds)
    iK, beta = pp.calculate_factorizations(X, Y, ell, sf2, sn2)
    dyn = lambda m, s: pp.predict_given_factorizations(X, ell, sf2, m, s, iK, beta)
    act = lambda m, s: pp.linear_action(W, b, m, s, True, maxa)
    radd = sp.objective_function(lambda m, s: pp.exponential_reward(m, s, Wr, tr),
                                 lambda m, s: sp.risk_of_collision(m, s, low, high), mu_obj)
    rmult = lambda m, s: sp.single_constraint(m, s, 1, low=0.3)
    Mn, Sn, tot = sp.safe_predict(m0, S0, H, lambda m, s: pp.propagate(m, s, act, dyn), radd, rmult, mu)

    def torch_total(Wt):
        iKt, bt = tp.calculate_factorizations(T(X), T(Y), T(ell), T(sf2), T(sn2))
        dyn_t = lambda m, s: tp.predict_given_factorizations(T(X), T(ell), T(sf2), m, s, iKt, bt)
        act_t = lambda m, s: tp.linear_action(Wt, T(b), m, s, T(maxa))
        radd_t = lambda m, s: tp.exponential_reward(m, s, T(Wr), T(tr)) - mu_obj * sp.box_risk_torch(m, s, (0, 2), low, high, 2.0, True)
        rmult_t = lambda m, s: sp.box_risk_torch(m, s, (1,), (0.3,), (float("inf"),), 1.0, True)
        return sp.safe_predict_torch(T(m0), T(S0), H, lambda m, s: tp.propagate(m, s, act_t, dyn_t), radd_t, rmult_t, mu)

    Wt = T(W).requires_grad_()
    Mt, St, tt = torch_total(Wt)
    assert abs(float(tot.item()) - float(tt.detach().item())) < 1e-10
    assert np.max(np.abs(Mn - Mt.detach().numpy())) < 1e-10 and np.max(np.abs(Sn - St.detach().numpy())) < 1e-10
    (gW,) = torch.autograd.grad(tt[0, 0], [Wt])
    eps = 1e-6
    for idx in [(0, 0), (0, 2)]:
        Wp, Wm = W.copy(), W.copy()
        Wp[idx] += eps; Wm[idx] -= eps
        fd = (float(torch_total(T(Wp))[2].item()) - float(torch_total(T(Wm))[2].item())) / (2 * eps)
        assert abs(fd - float(gW[idx])) < 1e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("N,M,D", [(60, 12, 3), (150, 25, 5)])
def test_fitc_staged_value_and_gradient(N, M, D):
    """oracle/fitc_staged.py (numpy statement of the FITC training objective + hand-derived adjoints, the algorithm
    a device SMGPR trainer implements) against torch autograd on the host training path's loss."""
    import torch
    from oracle import fitc_staged
    from pilco_b200 import gp_training
    rng = np.random.RandomState(N)
    X = rng.rand(N, D)
    y = np.sin(X).dot(rng.rand(D)) + 1e-2 * rng.randn(N)
    Z = rng.rand(M, D)
    ell, sf2, sn2 = 0.7 + rng.rand(D), 1.3, 0.05
    f, g = fitc_staged.fitc_nlml(X, y, Z, ell, sf2, sn2)
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True)
    tZ, tl, tf, tn = T(Z), T(ell), T(sf2), T(sn2)
    loss = gp_training.fitc_loss(torch.tensor(X), torch.tensor(y), tZ, tl, tf, tn)
    gZ, gl, gf, gn = torch.autograd.grad(loss, [tZ, tl, tf, tn])
    assert abs(f - float(loss.detach())) < 1e-9 * max(1.0, abs(f))
    rel = lambda a, b: np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(np.asarray(b))) + 1e-300)
    assert rel(g["ell"], gl.numpy()) < 1e-8 and rel(g["Z"], gZ.numpy()) < 1e-8
    assert rel(g["sf2"], gf.numpy()) < 1e-8 and rel(g["sn2"], gn.numpy()) < 1e-8


@pytest.mark.parametrize("n,D,E,mode", [(40, 3, 2, 0), (70, 5, 4, 0), (30, 3, 2, 1), (55, 12, 3, 0)])
def test_tape_formulation_equals_ordered_pair_vjp(n, D, E, mode):
    """oracle/staged.py: the tape-driven VJP (unordered pairs; row sums, column sums and H Z of every pair -- what the
    taped device forward leaves behind) gives the same (gm, gs) as the ordered-pair statement that torch autograd pins."""
    from oracle import staged as st, python_port as pp
    from util import make_gp_problem, make_input, scaled_err
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, E, seed=n + mode)
    if mode == 1:
        sf2, sn2 = np.ones(E), 1e-4 * np.ones(E)
    iK, beta = pp.calculate_factorizations(X, Y, ell, sf2, sn2)
    m, s = make_input(D, seed=20, scale=0.5)
    rng = np.random.RandomState(5)
    gM, gS, gV = rng.randn(E), rng.randn(E, E), rng.randn(D, E)
    ref = st.mm_backward_staged(X, ell, sf2, beta, iK, m[0], s, gM, gS, gV, mode)
    tape = st.mm_tape_forward(X, ell, sf2, beta, iK, m[0], s, mode)
    gm, gs = st.mm_backward_tape(X, ell, sf2, beta, m[0], s, gM, gS, gV, tape)
    assert scaled_err(gm, ref[0]) < 1e-10 and scaled_err(gs, ref[1]) < 1e-10
