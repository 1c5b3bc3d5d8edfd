"""numpy statement of the FITC training objective and its hand-derived gradient (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py) -- the algorithm a device kernel for ``SMGPR.optimize`` has to implement (SURVEY.md section 8f-1,
second half).  The reference trains ``gpflow.models.GPRFITC`` with TensorFlow autodiff
(pilco/models/smgpr.py:16-22, called from pilco/models/mgpr.py:47-75); the objective is GPflow 2.1's
``GPRFITC.maximum_log_likelihood_objective`` (not under /root/reference -- restated from its published form, the same
restatement the host path ``pilco_b200.gp_training.fitc_loss`` uses), no priors, trainable inducing inputs Z.

Every stage is an explicit matrix operation with its adjoint, in the order a device implementation runs them:
  Gram blocks -> chol(Kuu) -> V = Luu^-1 Kuf -> nu -> B = I + V diag(1/nu) V^T -> chol(B) -> alpha, gamma -> value
and the reverse sweep back to (ell, sf2, sn2, Z).  Checked against torch autograd in tests/test_oracle.py.
PARITY UNPINNED by the reference (no reference test looks at training trajectories or gradients).
"""
import math

import numpy as np
from scipy.linalg import solve_triangular


def _gram(A, B, ell, sf2):
    a, b = A / ell, B / ell
    d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T
    return sf2 * np.exp(-0.5 * np.maximum(d2, 0.0))


def _chol_backward(L, Lbar):
    """adjoint of L = chol(K) (lower): Kbar = sym( L^-T Phi(L^T Lbar) L^-1 ), Phi = lower triangle, diagonal halved."""
    P = np.tril(L.T @ np.tril(Lbar))
    P[np.diag_indices_from(P)] *= 0.5
    S = solve_triangular(L, solve_triangular(L, P.T, lower=True, trans="T").T, lower=True, trans="T")   # L^-T P L^-1
    return 0.5 * (S + S.T)


def fitc_nlml(X, y, Z, ell, sf2, sn2, jitter=1e-6, grad=True):
    """-> nlml (float) and, with ``grad``, d nlml / d (ell [D], sf2, sn2, Z [M,D])."""
    N, M = X.shape[0], Z.shape[0]
    Kuf = _gram(Z, X, ell, sf2)                              # [M,N]
    Kuu0 = _gram(Z, Z, ell, sf2)
    Luu = np.linalg.cholesky(Kuu0 + jitter * np.eye(M))
    V = solve_triangular(Luu, Kuf, lower=True)               # [M,N]
    nu = sf2 + sn2 - (V * V).sum(0)                          # [N]
    W = V / nu
    B = np.eye(M) + W @ V.T
    L = np.linalg.cholesky(B)
    alpha = V @ (y / nu)
    gamma = solve_triangular(L, alpha, lower=True)
    f = 0.5 * (y * y / nu).sum() - 0.5 * gamma @ gamma + 0.5 * np.log(nu).sum() + np.log(np.diag(L)).sum() \
        + 0.5 * N * math.log(2.0 * math.pi)
    if not grad:
        return f
    # ---- reverse sweep ---------------------------------------------------------------------------
    gamma_b = -gamma
    alpha_b = solve_triangular(L, gamma_b, lower=True, trans="T")           # L^-T gamma_bar
    L_b = -np.outer(alpha_b, gamma) + np.diag(1.0 / np.diag(L))
    B_b = _chol_backward(L, L_b)                                            # symmetric
    BV = B_b @ V
    V_b = 2.0 * BV / nu + np.outer(alpha_b, y / nu)
    nu_b = -(V * BV).sum(0) / nu ** 2 - (alpha_b @ V) * y / nu ** 2 - 0.5 * y * y / nu ** 2 + 0.5 / nu
    sf2_b = nu_b.sum()
    sn2_b = nu_b.sum()
    V_b = V_b - 2.0 * V * nu_b
    Kuf_b = solve_triangular(Luu, V_b, lower=True, trans="T")               # Luu^-T V_bar
    Luu_b = -Kuf_b @ V.T
    Kuu_b = _chol_backward(Luu, Luu_b)                                      # symmetric
    # ---- kernel adjoints: K = sf2 exp(-0.5 sum_d (a_d - b_d)^2 / ell_d^2) ----------------------------
    Guf = Kuf_b * Kuf                                                       # [M,N]
    Guu = Kuu_b * Kuu0                                                      # [M,M] (jitter carries no parameters)
    sf2_b += (Guf.sum() + Guu.sum()) / sf2
    il2 = 1.0 / ell ** 2
    ell_b = np.zeros_like(ell)
    Z_b = np.zeros_like(Z)
    for d in range(X.shape[1]):
        dzx = Z[:, d][:, None] - X[:, d][None, :]                           # z_i - x_n
        dzz = Z[:, d][:, None] - Z[:, d][None, :]                           # z_i - z_j
        ell_b[d] = ((Guf * dzx ** 2).sum() + (Guu * dzz ** 2).sum()) / ell[d] ** 3
        # d/dz_i of -0.5 (z_i - x_n)^2/ell^2 = -(z_i - x_n)/ell^2 ; Kuu: both index positions of z_i
        Z_b[:, d] = -(Guf * dzx).sum(1) * il2[d] - ((Guu + Guu.T) * dzz).sum(1) * il2[d]
    return f, dict(ell=ell_b, sf2=sf2_b, sn2=sn2_b, Z=Z_b)
