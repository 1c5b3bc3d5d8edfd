"""GP hyper-parameter training (``MGPR.optimize`` / ``PILCO.optimize_models``), HOST logic.

Scope note (DESIGN.md "Out of scope / next"): hyper-parameter training is the
step *before* the hot path and is marked "next" in the scope table.  It is implemented here on the
host with torch-CPU autograd + SciPy L-BFGS-B (the optimiser the reference uses through GPflow,
pilco/models/mgpr.py:47-75); no reference test pins its trajectory -- every parity test feeds whatever
hyper-parameters come out of training to the oracle.  None of this code is on the moment-matching path.

Losses (GPflow 2.1 semantics):
  GPR      -log p(y|X,theta) - sum log prior      (mgpr.py:28-36; Gamma priors on ell and sf2)
  GPRFITC  FITC bound with trainable inducing inputs Z  (smgpr.py:16-22; no priors)
"""
import math

import numpy as np
import scipy.optimize
import torch

from .params import Parameter

F64 = torch.float64


def _softplus_fwd(theta, lower):
    return lower + torch.nn.functional.softplus(theta)


def minimize(loss_fn, params, maxiter=None):
    """L-BFGS-B over the trainable ``params`` (list of Parameter).  ``loss_fn(values)`` receives a list of
    constrained torch tensors (one per entry of ``params``, trainable or not) and returns a scalar."""
    train = [p for p in params if p.trainable]
    if not train:
        return float(loss_fn([torch.as_tensor(p.value(), dtype=F64) for p in params]))
    sizes = [int(np.prod(p.shape)) if p.shape else 1 for p in train]
    x0 = np.concatenate([np.asarray(p.unconstrained, dtype=np.float64).ravel() for p in train])

    def unpack(x):
        vals, off, leaves = [], 0, []
        it = iter(range(len(train)))
        for p in params:
            if p.trainable:
                k = next(it)
                th = torch.tensor(x[off:off + sizes[k]].reshape(p.shape), dtype=F64, requires_grad=True)
                off += sizes[k]
                leaves.append(th)
                lower = getattr(p.transform, "lower", None)
                vals.append(th if lower is None else _softplus_fwd(th, lower))
            else:
                vals.append(torch.as_tensor(p.value(), dtype=F64))
        return vals, leaves

    def fun(x):
        vals, leaves = unpack(x)
        try:
            loss = loss_fn(vals)
            grads = torch.autograd.grad(loss, leaves)
            f = float(loss.detach())
            g = np.concatenate([gr.detach().numpy().ravel() for gr in grads])
        except (RuntimeError, torch.linalg.LinAlgError):
            return 1e100, np.zeros_like(x)
        if not np.isfinite(f) or not np.all(np.isfinite(g)):
            return 1e100, np.zeros_like(x)
        return f, g

    opts = {} if maxiter is None else {"maxiter": int(maxiter)}
    res = scipy.optimize.minimize(fun, x0, jac=True, method="L-BFGS-B", options=opts)
    off = 0
    for p, k in zip(train, sizes):
        p.set_unconstrained(res.x[off:off + k].reshape(p.shape))
        off += k
    return float(res.fun)


def _se_ard(X1, X2, ell, sf2):
    a, b = X1 / ell, X2 / ell
    d2 = (a * a).sum(-1)[:, None] + (b * b).sum(-1)[None, :] - 2.0 * a @ b.T
    return sf2 * torch.exp(-0.5 * d2.clamp_min(0.0))


def _gamma_logpdf(x, alpha, rate):
    return alpha * math.log(rate) - math.lgamma(alpha) + (alpha - 1.0) * torch.log(x) - rate * x


def gpr_loss(X, y, ell, sf2, sn2, ell_prior=None, sf2_prior=None):
    n = X.shape[0]
    K = _se_ard(X, X, ell, sf2) + sn2 * torch.eye(n, dtype=F64)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve(y[:, None], L)[:, 0]
    ll = -0.5 * (y * alpha).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2.0 * math.pi)
    if ell_prior is not None:
        ll = ll + _gamma_logpdf(ell, *ell_prior).sum()
    if sf2_prior is not None:
        ll = ll + _gamma_logpdf(sf2, *sf2_prior).sum()
    return -ll


def fitc_loss(X, y, Z, ell, sf2, sn2, jitter=1e-6):
    n, M = X.shape[0], Z.shape[0]
    Kuf = _se_ard(Z, X, ell, sf2)
    Kuu = _se_ard(Z, Z, ell, sf2) + jitter * torch.eye(M, dtype=F64)
    Luu = torch.linalg.cholesky(Kuu)
    V = torch.linalg.solve_triangular(Luu, Kuf, upper=False)
    nu = sf2 - (V * V).sum(0) + sn2
    B = torch.eye(M, dtype=F64) + (V / nu) @ V.T
    L = torch.linalg.cholesky(B)
    beta = y / nu
    alpha = V @ beta
    gamma = torch.linalg.solve_triangular(L, alpha[:, None], upper=False)[:, 0]
    maha = -0.5 * (y * y / nu).sum() + 0.5 * (gamma * gamma).sum()
    logdet = -0.5 * torch.log(nu).sum() - torch.log(torch.diagonal(L)).sum()
    return -(maha + logdet - 0.5 * n * math.log(2.0 * math.pi))
