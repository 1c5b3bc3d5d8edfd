"""GPU tests of the device GP training objective (pilco_gp_nlml): value and gradient against
torch autograd on the host objective (gp_training.gpr_loss, GPflow 2.1 semantics), rtol 1e-8; and the lock-step
driver wired into MGPR.optimize."""
import numpy as np
import pytest
import torch

from util import scaled_err, make_gp_problem

pytestmark = pytest.mark.gpu
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)


@pytest.mark.parametrize("n,D,E,B", [(60, 3, 2, 1), (130, 5, 3, 2), (300, 12, 2, 1)])
def test_gp_nlml_matches_autograd(n, D, E, B):
    from pilco_b200 import engine, gp_training
    X, Y, ell, sf2, sn2 = make_gp_problem(n, D, E, seed=n, noise=1e-2)
    rng = np.random.RandomState(1)
    ellB = np.stack([ell * (1 + 0.1 * rng.rand(E, D)) for _ in range(B)])
    sf2B = np.stack([sf2 * (1 + 0.1 * rng.rand(E)) for _ in range(B)])
    sn2B = np.stack([sn2 * (1 + 0.1 * rng.rand(E)) for _ in range(B)])
    ev = engine.GpNlml(X, Y, B)
    nlml, g_ell, g_sf2, g_sn2, bad = ev(ellB, sf2B, sn2B)
    assert not bad.any()
    for b in range(B):
        for e in range(E):
            tl, tv, tn = T(ellB[b, e]).requires_grad_(), T(sf2B[b, e]).requires_grad_(), T(sn2B[b, e]).requires_grad_()
            loss = gp_training.gpr_loss(T(X), T(Y[:, e]), tl, tv, tn)
            gl, gv, gn = torch.autograd.grad(loss, [tl, tv, tn])
            assert abs(nlml[b, e] - loss.item()) < 1e-8 * abs(loss.item())
            assert scaled_err(g_ell[b, e], gl.numpy()) < 1e-7
            assert abs(g_sf2[b, e] - gv.item()) < 1e-7 * max(1.0, abs(gv.item()))
            assert abs(g_sn2[b, e] - gn.item()) < 1e-7 * max(1.0, abs(gn.item()))


def test_mgpr_optimize_on_device():
    from pilco.models import MGPR
    from gpflow import set_trainable
    np.random.seed(0)
    X = np.random.rand(80, 3)
    Y = np.sin(3 * X).dot(np.random.rand(3, 2)) + 1e-2 * np.random.randn(80, 2)
    m = MGPR((X, Y))
    l0 = [mod.training_loss() for mod in m.models]
    m.optimize(restarts=2)
    l1 = [mod.training_loss() for mod in m.models]
    assert all(b < a for a, b in zip(l0, l1))
    assert np.all(m.lengthscales > 0) and np.all(m.noise >= 1e-6) and np.all(np.isfinite(m.variance))
    # a frozen noise stays where it was put (inv_double_pendulum.py:83-85 pattern)
    m.models[0].likelihood.variance.assign(0.01)
    set_trainable(m.models[0].likelihood.variance, False)
    m.optimize(restarts=0)
    assert abs(m.noise[0] - 0.01) < 1e-12
    # device optimum is a stationary point of the host objective as well
    m2 = MGPR((X, Y))
    m2.optimize_host(restarts=0)
    m3 = MGPR((X, Y))
    m3.optimize(restarts=0)
    for a, b in zip(m2.models, m3.models):
        assert abs(a.training_loss() - b.training_loss()) < 1e-4 * max(1.0, abs(a.training_loss()))


@pytest.mark.parametrize("N,M,D,E,B", [(300, 50, 4, 2, 2), (2000, 200, 12, 2, 1)])
def test_fitc_nlml_matches_autograd(N, M, D, E, B):
    """pilco_fitc_nlml (sparse half): the FITC bound and its gradient w.r.t. lengthscales, variances
    and EVERY inducing input against torch autograd on the host objective (gp_training.fitc_loss, GPflow 2.1's
    GPRFITC semantics as SMGPR.optimize minimises it, smgpr.py:16-22) -- including BASELINE.json's N=2000, M=200."""
    from pilco_b200 import engine, gp_training
    X, Y, ell, sf2, sn2 = make_gp_problem(N, D, E, seed=N, noise=1e-2)
    rng = np.random.RandomState(2)
    ZB = rng.rand(B, E, M, D)
    ellB = np.stack([ell * (1 + 0.1 * rng.rand(E, D)) for _ in range(B)])
    sf2B = np.stack([sf2 * (1 + 0.1 * rng.rand(E)) for _ in range(B)])
    sn2B = np.stack([sn2 * (1 + 0.1 * rng.rand(E)) for _ in range(B)])
    ev = engine.FitcNlml(X, Y, M, B)
    nlml, g_ell, g_sf2, g_sn2, g_Z, bad = ev(ZB, ellB, sf2B, sn2B)
    assert not bad.any()
    for b in range(B):
        for e in range(E):
            tz = T(ZB[b, e]).requires_grad_()
            tl, tv, tn = T(ellB[b, e]).requires_grad_(), T(sf2B[b, e]).requires_grad_(), T(sn2B[b, e]).requires_grad_()
            loss = gp_training.fitc_loss(T(X), T(Y[:, e]), tz, tl, tv, tn)
            gz, gl, gv, gn = torch.autograd.grad(loss, [tz, tl, tv, tn])
            assert abs(nlml[b, e] - loss.item()) < 1e-8 * abs(loss.item())
            assert scaled_err(g_ell[b, e], gl.numpy()) < 1e-7
            assert scaled_err(g_Z[b, e], gz.numpy()) < 1e-7
            assert abs(g_sf2[b, e] - gv.item()) < 1e-7 * max(1.0, abs(gv.item()))
            assert abs(g_sn2[b, e] - gn.item()) < 1e-7 * max(1.0, abs(gn.item()))


def test_smgpr_optimize_on_device():
    """SMGPR.optimize runs on the device (no torch-CPU autograd): the loss goes down for every output, inducing inputs
    move, and the optimum is as good as the host path's."""
    from pilco.models import SMGPR
    np.random.seed(0)
    X = np.random.rand(150, 3)
    Y = np.sin(3 * X).dot(np.random.rand(3, 2)) + 1e-2 * np.random.randn(150, 2)
    m = SMGPR((X, Y), num_induced_points=20)
    Z0 = np.asarray(m.Z).copy()
    l0 = [mod.training_loss() for mod in m.models]
    m.optimize(restarts=1, maxiter=150)
    l1 = [mod.training_loss() for mod in m.models]
    assert all(b < a for a, b in zip(l0, l1))
    assert not np.allclose(np.asarray(m.Z), Z0)
    np.random.seed(0)
    mh = SMGPR((X, Y), num_induced_points=20)
    mh.optimize_host(restarts=0, maxiter=150)
    for a, b in zip(mh.models, m.models):
        assert b.training_loss() < a.training_loss() + 1e-2 * max(1.0, abs(a.training_loss()))
    M_, S_, V_ = m.predict_on_noisy_inputs(np.random.rand(1, 3), 0.01 * np.eye(3))
    assert np.all(np.isfinite(M_)) and np.all(np.isfinite(S_))
