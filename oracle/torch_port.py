"""Vectorised torch (fp64) restatement of the reference's Python hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Same op-for-op structure as ``python_port`` (materialised [E,E,N,N] tensors, like TensorFlow), but in
torch so that (1) autograd provides the gradient oracle for the hand-derived device VJPs and
(2) it runs multi-threaded on the host cores as the timed "reference-equivalent CPU path" of
``bench.py`` (TensorFlow/GPflow are not installable offline).
Citations are ``file:line`` relative to ``/root/reference/``.
"""
import torch

F64 = torch.float64


def se_ard_K(X1, X2, ell, sf2):
    """gpflow SquaredExponential.K per output (mgpr.py:154-157)."""
    a = X1[None] / ell[:, None, :]
    b = X2[None] / ell[:, None, :]
    d2 = (a * a).sum(-1)[:, :, None] + (b * b).sum(-1)[:, None, :] - 2.0 * a @ b.transpose(1, 2)
    return sf2[:, None, None] * torch.exp(-0.5 * d2.clamp_min(0.0))


def calculate_factorizations(X, Y, ell, sf2, sn2):
    """mgpr.py:81-89"""
    E, N = Y.shape[1], X.shape[0]
    K = se_ard_K(X, X, ell, sf2)
    eye = torch.eye(N, dtype=F64).expand(E, N, N)
    L = torch.linalg.cholesky(K + sn2[:, None, None] * eye)
    iK = torch.cholesky_solve(eye, L)
    beta = torch.cholesky_solve(Y.T[:, :, None], L)[:, :, 0]
    return iK, beta


def fitc_factorizations(X, Z, Y, ell, sf2, sn2):
    """smgpr.py:24-45 -> iK [E,M,M], beta [E,M] over the inducing points Z (batched over the outputs)."""
    E, Mi = Y.shape[1], Z.shape[0]
    eye = torch.eye(Mi, dtype=F64).expand(E, Mi, Mi)
    Kmm = se_ard_K(Z, Z, ell, sf2) + 1e-6 * eye
    Kmn = se_ard_K(Z, X, ell, sf2)
    L = torch.linalg.cholesky(Kmm)
    V = torch.linalg.solve_triangular(L, Kmn, upper=False)
    G = torch.sqrt(1.0 + (sf2[:, None] - (V ** 2).sum(1)) / sn2[:, None])
    V = V / G[:, None, :]
    Am = torch.linalg.cholesky(V @ V.transpose(1, 2) + sn2[:, None, None] * eye)
    At = L @ Am
    iAt = torch.linalg.solve_triangular(At, eye, upper=False)
    rhs = ((V / G[:, None, :]) @ Y.T[:, :, None])
    tmp = torch.cholesky_solve(rhs, Am)
    beta = torch.linalg.solve_triangular(L.transpose(1, 2), tmp, upper=True)[:, :, 0]
    iK = torch.cholesky_solve(eye, L) - sn2[:, None, None] * (iAt.transpose(1, 2) @ iAt)
    return iK, beta


def predict_given_factorizations(C, ell, sf2, m, s, iK, beta):
    """mgpr.py:91-149; C = centres."""
    E, D = ell.shape
    N = C.shape[0]
    s4 = s[None, None].expand(E, E, D, D)
    inp = (C - m)[None].expand(E, N, D)
    iL = torch.diag_embed(1.0 / ell)
    iN = inp @ iL
    B = iL @ s4[0] @ iL + torch.eye(D, dtype=F64)
    t = torch.linalg.solve(B.transpose(1, 2), iN.transpose(1, 2)).transpose(1, 2)
    lb = torch.exp(-(iN * t).sum(-1) / 2.0) * beta
    tiL = t @ iL
    c = sf2 / torch.sqrt(torch.linalg.det(B))
    M = (lb.sum(-1) * c)[:, None]
    V = (tiL.transpose(1, 2) @ lb[:, :, None])[..., 0] * c[:, None]

    Rm = s4 @ torch.diag_embed(1.0 / ell[None, :, :] ** 2 + 1.0 / ell[:, None, :] ** 2) + torch.eye(D, dtype=F64)
    Xa = inp[None, :, :, :] / (ell ** 2)[:, None, None, :]
    X2 = -inp[:, None, :, :] / (ell ** 2)[None, :, None, :]
    Q = torch.linalg.solve(Rm, s4) / 2.0
    Xs = ((Xa @ Q) * Xa).sum(-1)
    X2s = ((X2 @ Q) * X2).sum(-1)
    maha = -2.0 * ((Xa @ Q) @ X2.transpose(2, 3)) + Xs[:, :, :, None] + X2s[:, :, None, :]
    k = torch.log(sf2)[:, None] - (iN ** 2).sum(-1) / 2.0
    L = torch.exp(k[:, None, :, None] + k[None, :, None, :] + maha)
    S = (beta[:, None, None, :].expand(E, E, 1, N) @ L @ beta[None, :, :, None].expand(E, E, N, 1))[:, :, 0, 0]
    diagL = torch.stack([L[e, e] for e in range(E)])
    if iK is not None:
        S = S - torch.diag((iK * diagL).sum((1, 2)))
    S = S / torch.sqrt(torch.linalg.det(Rm))
    S = S + torch.diag(sf2)
    S = S - M @ M.T
    return M.T, S, V.T


def squash_sin(m, s, max_action):
    """controllers.py:13-36; max_action [1,k] tensor."""
    k = m.shape[1]
    ds = torch.diagonal(s)
    M = max_action * torch.exp(-ds / 2.0) * torch.sin(m)
    lq = -(ds[:, None] + ds[None, :]) / 2.0
    q = torch.exp(lq)
    S = (torch.exp(lq + s) - q) * torch.cos(m.T - m) - (torch.exp(lq - s) - q) * torch.cos(m.T + m)
    S = max_action * max_action.T * S / 2.0
    C = max_action * torch.diag(torch.exp(-ds / 2.0) * torch.cos(m)[0])
    return M, S, C.reshape(k, k)


def linear_action(W, b, m, s, max_action=None):
    """controllers.py:46-58 (squash when max_action is given)."""
    M = m @ W.T + b
    S = W @ s @ W.T
    V = W.T
    if max_action is not None:
        M, S, V2 = squash_sin(M, S, max_action)
        V = V @ V2
    return M, S, V


def rbf_action(Xc, Yc, ell, m, s, max_action=None):
    """controllers.py:108-121 (sf2=1, sn2=1e-4 fixed)."""
    U = Yc.shape[1]
    sf2 = torch.ones(U, dtype=F64)
    sn2 = 1e-4 * torch.ones(U, dtype=F64)
    _, beta = calculate_factorizations(Xc, Yc, ell, sf2, sn2)
    M, S, V = predict_given_factorizations(Xc, ell, sf2, m, s, None, beta)
    S = S - torch.diag(sf2 - 1e-6)
    if max_action is not None:
        M, S, V2 = squash_sin(M, S, max_action)
        V = V @ V2
    return M, S, V


def exponential_reward(m, s, W, t):
    """rewards.py:19-51 (mean only; the loss uses [0], pilco.py:133)."""
    D = m.shape[1]
    SW = s @ W
    iSpW = torch.linalg.solve((torch.eye(D, dtype=F64) + SW).T, W.T).T
    muR = torch.exp(-(m - t) @ iSpW @ (m - t).T / 2.0) / torch.sqrt(torch.linalg.det(torch.eye(D, dtype=F64) + SW))
    return muR.reshape(1, 1)


def propagate(m_x, s_x, action_fn, dynamics_fn):
    """pilco.py:138-153"""
    m_u, s_u, c_xu = action_fn(m_x, s_x)
    m = torch.cat([m_x, m_u], dim=1)
    s1 = torch.cat([s_x, s_x @ c_xu], dim=1)
    s2 = torch.cat([(s_x @ c_xu).T, s_u], dim=1)
    s = torch.cat([s1, s2], dim=0)
    M_dx, S_dx, C_dx = dynamics_fn(m, s)
    M_x = M_dx + m_x
    S_x = S_dx + s_x + s1 @ C_dx + C_dx.T @ s1.T
    return M_x, S_x


def predict(m_x, s_x, n, action_fn, dynamics_fn, reward_fn):
    """pilco.py:118-136"""
    total = torch.zeros((1, 1), dtype=F64)
    for _ in range(n):
        r = reward_fn(m_x, s_x)
        m_x, s_x = propagate(m_x, s_x, action_fn, dynamics_fn)
        total = total + r
    return m_x, s_x, total
