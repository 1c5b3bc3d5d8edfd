"""GP hyper-parameter training on the device (replaces the per-model
``gpflow.optimizers.Scipy().minimize(model.training_loss, ...)`` loop of ``MGPR.optimize``,
pilco/models/mgpr.py:47-75).

All E outputs x (1 + restarts) initialisations are optimised in lock step: every L-BFGS-B evaluation of every
problem is served by ONE batched ``pilco_gp_nlml`` call (Gram, blocked Cholesky, inverse, NLML and its analytic
gradient on the device).  The host adds the Gamma log-priors of mgpr.py:33-34 and the softplus chain rule.
Restart k >= 1 starts from ``randomize`` (mgpr.py:8-15); the best final loss per output wins (the reference's
bookkeeping keeps the last restart -- see the note in ``MGPR.optimize``)."""
import numpy as np

from . import engine
from .policy_opt import LockstepLBFGS, BIG


def _sigmoid(t):
    return 1.0 / (1.0 + np.exp(-t))


def optimize_mgpr(mgpr, restarts=1, maxiter=None):
    models = mgpr.models
    E, D = len(models), mgpr.num_dims
    B = int(restarts) + 1
    X, Y = mgpr.X, mgpr.Y
    params = [[m.kernel.lengthscales, m.kernel.variance, m.likelihood.variance] for m in models]
    lowers = np.array([[p.transform.lower for p in ps] for ps in params])          # [E,3]
    train = np.array([[p.trainable for p in ps] for ps in params])                 # [E,3]
    pri_l = [ps[0].prior for ps in params]
    pri_v = [ps[1].prior for ps in params]
    P = D + 2
    # initial constrained values [B,E,...]
    ell0 = np.empty((B, E, D)); sf20 = np.empty((B, E)); sn20 = np.empty((B, E))
    for e, ps in enumerate(params):
        ell0[0, e], sf20[0, e], sn20[0, e] = np.asarray(ps[0]), float(ps[1]), float(ps[2])
        for k in range(1, B):                                                       # randomize(): mgpr.py:8-15
            ell0[k, e] = 1.0 + 0.01 * np.random.normal(size=D) if train[e, 0] else ell0[0, e]
            sf20[k, e] = 1.0 + 0.01 * np.random.normal() if train[e, 1] else sf20[0, e]
            sn20[k, e] = 1.0 + 0.01 * np.random.normal() if train[e, 2] else sn20[0, e]

    def inv_softplus(x, lower):
        y = np.maximum(np.asarray(x, dtype=np.float64) - lower, 1e-300)
        return y + np.log(-np.expm1(-y))

    x0 = np.empty((B, E, P))
    for e in range(E):
        x0[:, e, :D] = inv_softplus(ell0[:, e], lowers[e, 0])
        x0[:, e, D] = inv_softplus(sf20[:, e], lowers[e, 1])
        x0[:, e, D + 1] = inv_softplus(sn20[:, e], lowers[e, 2])
    dev_eval = engine.GpNlml(X, Y, B)
    low_l = lowers[None, :, 0, None]; low_v = lowers[None, :, 1]; low_n = lowers[None, :, 2]

    def evaluate(xs):
        th = xs.reshape(B, E, P)
        ell = low_l + np.logaddexp(0.0, th[:, :, :D])
        sf2 = low_v + np.logaddexp(0.0, th[:, :, D])
        sn2 = low_n + np.logaddexp(0.0, th[:, :, D + 1])
        nlml, g_ell, g_sf2, g_sn2, bad = dev_eval(ell, sf2, sn2)
        loss = nlml.copy()
        for e in range(E):                                                          # Gamma log-priors (alpha, rate)
            if pri_l[e] is not None:
                al, rt = pri_l[e]
                loss[:, e] -= ((al - 1.0) * np.log(ell[:, e]) - rt * ell[:, e]).sum(-1)
                g_ell[:, e] -= (al - 1.0) / ell[:, e] - rt
            if pri_v[e] is not None:
                al, rt = pri_v[e]
                loss[:, e] -= (al - 1.0) * np.log(sf2[:, e]) - rt * sf2[:, e]
                g_sf2[:, e] -= (al - 1.0) / sf2[:, e] - rt
        g = np.empty((B, E, P))
        g[:, :, :D] = g_ell * _sigmoid(th[:, :, :D]) * train[None, :, 0, None]
        g[:, :, D] = g_sf2 * _sigmoid(th[:, :, D]) * train[None, :, 1]
        g[:, :, D + 1] = g_sn2 * _sigmoid(th[:, :, D + 1]) * train[None, :, 2]
        badm = bad[:, None] | ~np.isfinite(loss) | ~np.isfinite(g).all(-1)
        loss[badm] = BIG
        g[badm] = 0.0
        return loss.reshape(-1), g.reshape(B * E, P)

    finals = LockstepLBFGS(evaluate, x0.reshape(B * E, P), 15000 if maxiter is None else int(maxiter)).run()
    losses = np.array([f[0] for f in finals]).reshape(B, E)
    xs = np.stack([f[1] for f in finals]).reshape(B, E, P)
    for e, ps in enumerate(params):
        k = int(np.argmin(losses[:, e]))
        if train[e, 0]:
            ps[0].set_unconstrained(xs[k, e, :D])
        if train[e, 1]:
            ps[1].set_unconstrained(xs[k, e, D])
        if train[e, 2]:
            ps[2].set_unconstrained(xs[k, e, D + 1])
    return losses.min(axis=0)


def optimize_smgpr(smgpr, restarts=1, maxiter=None):
    """FITC training with trainable inducing inputs on the device (replaces the GPRFITC + TF-autodiff loop of
    ``SMGPR.optimize``, pilco/models/smgpr.py:16-22 via mgpr.py:47-75): all E outputs x (1 + restarts) initialisations
    advance in lock step, every L-BFGS-B evaluation is ONE batched ``pilco_fitc_nlml`` call (value and gradient w.r.t.
    lengthscales, signal / noise variance and every inducing input).  No priors (smgpr.py sets none)."""
    models = smgpr.models
    E, D, Mi = len(models), smgpr.num_dims, int(smgpr.num_induced_points)
    B = int(restarts) + 1
    params = [[m.kernel.lengthscales, m.kernel.variance, m.likelihood.variance, m.inducing_variable.Z] for m in models]
    lowers = np.array([[p.transform.lower for p in ps[:3]] for ps in params])      # [E,3]
    train = np.array([[p.trainable for p in ps] for ps in params])                 # [E,4]
    P = D + 2 + Mi * D
    ell0 = np.empty((B, E, D)); sf20 = np.empty((B, E)); sn20 = np.empty((B, E)); Z0 = np.empty((B, E, Mi, D))
    for e, ps in enumerate(params):
        ell0[0, e], sf20[0, e], sn20[0, e] = np.asarray(ps[0]), float(ps[1]), float(ps[2])
        Z0[:, e] = np.asarray(ps[3])
        for k in range(1, B):                                                       # randomize(): mgpr.py:8-15
            ell0[k, e] = 1.0 + 0.01 * np.random.normal(size=D) if train[e, 0] else ell0[0, e]
            sf20[k, e] = 1.0 + 0.01 * np.random.normal() if train[e, 1] else sf20[0, e]
            sn20[k, e] = 1.0 + 0.01 * np.random.normal() if train[e, 2] else sn20[0, e]

    def inv_softplus(x, lower):
        y = np.maximum(np.asarray(x, dtype=np.float64) - lower, 1e-300)
        return y + np.log(-np.expm1(-y))

    x0 = np.empty((B, E, P))
    for e in range(E):
        x0[:, e, :D] = inv_softplus(ell0[:, e], lowers[e, 0])
        x0[:, e, D] = inv_softplus(sf20[:, e], lowers[e, 1])
        x0[:, e, D + 1] = inv_softplus(sn20[:, e], lowers[e, 2])
        x0[:, e, D + 2:] = Z0[:, e].reshape(B, -1)
    dev_eval = engine.FitcNlml(smgpr.X, smgpr.Y, Mi, B)
    low_l = lowers[None, :, 0, None]; low_v = lowers[None, :, 1]; low_n = lowers[None, :, 2]

    def evaluate(xs):
        th = xs.reshape(B, E, P)
        ell = low_l + np.logaddexp(0.0, th[:, :, :D])
        sf2 = low_v + np.logaddexp(0.0, th[:, :, D])
        sn2 = low_n + np.logaddexp(0.0, th[:, :, D + 1])
        Z = th[:, :, D + 2:].reshape(B, E, Mi, D)
        loss, g_ell, g_sf2, g_sn2, g_Z, bad = dev_eval(Z, ell, sf2, sn2)
        g = np.empty((B, E, P))
        g[:, :, :D] = g_ell * _sigmoid(th[:, :, :D]) * train[None, :, 0, None]
        g[:, :, D] = g_sf2 * _sigmoid(th[:, :, D]) * train[None, :, 1]
        g[:, :, D + 1] = g_sn2 * _sigmoid(th[:, :, D + 1]) * train[None, :, 2]
        g[:, :, D + 2:] = g_Z.reshape(B, E, -1) * train[None, :, 3, None]
        loss = loss.copy()
        badm = bad[:, None] | ~np.isfinite(loss) | ~np.isfinite(g).all(-1)
        loss[badm] = BIG
        g[badm] = 0.0
        return loss.reshape(-1), g.reshape(B * E, P)

    finals = LockstepLBFGS(evaluate, x0.reshape(B * E, P), 15000 if maxiter is None else int(maxiter)).run()
    losses = np.array([f[0] for f in finals]).reshape(B, E)
    xs = np.stack([f[1] for f in finals]).reshape(B, E, P)
    for e, ps in enumerate(params):
        k = int(np.argmin(losses[:, e]))
        if train[e, 0]:
            ps[0].set_unconstrained(xs[k, e, :D])
        if train[e, 1]:
            ps[1].set_unconstrained(xs[k, e, D])
        if train[e, 2]:
            ps[2].set_unconstrained(xs[k, e, D + 1])
        if train[e, 3]:
            ps[3].assign(xs[k, e, D + 2:].reshape(Mi, D))
    return losses.min(axis=0)
