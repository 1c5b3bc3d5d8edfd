"""Vectorised numpy (fp64) restatement of the reference's Python hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Follows the TensorFlow code op for op, including the materialised ``[E,E,N,N]`` tensors, so it
also serves as the "reference-equivalent CPU path" whose cost structure matches the reference.
Citations are ``file:line`` relative to ``/root/reference/``.

Conventions follow the Python reference: ``m`` is ``[1, D]``, ``s`` is ``[D, D]``; outputs are
``M[1,E], S[E,E], V[D,E]``.
"""
import numpy as np


def se_ard_K(X1, X2, ell, sf2):
    """gpflow SquaredExponential.K as used at pilco/models/mgpr.py:154-157 (per output e)."""
    a = X1[None, :, :] / ell[:, None, :]
    b = X2[None, :, :] / ell[:, None, :]
    d2 = (a * a).sum(-1)[:, :, None] + (b * b).sum(-1)[:, None, :] - 2.0 * a @ b.transpose(0, 2, 1)
    return sf2[:, None, None] * np.exp(-0.5 * np.maximum(d2, 0.0))


def calculate_factorizations(X, Y, ell, sf2, sn2):
    """pilco/models/mgpr.py:81-89 -> iK[E,N,N], beta[E,N]."""
    E = Y.shape[1]
    N = X.shape[0]
    K = se_ard_K(X, X, ell, sf2)
    eye = np.broadcast_to(np.eye(N), (E, N, N))
    L = np.linalg.cholesky(K + sn2[:, None, None] * eye)
    iK = np.stack([np.linalg.solve(L[e].T, np.linalg.solve(L[e], np.eye(N))) for e in range(E)])
    beta = np.stack([np.linalg.solve(L[e].T, np.linalg.solve(L[e], Y[:, e])) for e in range(E)])
    return iK, beta


def fitc_factorizations(X, Z, Y, ell, sf2, sn2):
    """pilco/models/smgpr.py:24-45 -> iK[E,M,M], beta[E,M] over the inducing points Z."""
    E = Y.shape[1]
    Mi = Z.shape[0]
    eye = np.eye(Mi)
    iK = np.zeros((E, Mi, Mi))
    beta = np.zeros((E, Mi))
    Kmm_all = se_ard_K(Z, Z, ell, sf2) + 1e-6 * eye[None]
    Kmn_all = se_ard_K(Z, X, ell, sf2)
    for e in range(E):
        Kmm, Kmn = Kmm_all[e], Kmn_all[e]
        L = np.linalg.cholesky(Kmm)
        V = np.linalg.solve(L, Kmn)
        G = sf2[e] - (V ** 2).sum(0)
        G = np.sqrt(1.0 + G / sn2[e])
        V = V / G[None, :]
        Am = np.linalg.cholesky(V @ V.T + sn2[e] * eye)
        At = L @ Am
        iAt = np.linalg.solve(At, eye)
        rhs = (V / G[None, :]) @ Y[:, e]
        tmp = np.linalg.solve(Am.T, np.linalg.solve(Am, rhs))
        beta[e] = np.linalg.solve(L.T, tmp)
        iB = iAt.T @ iAt * sn2[e]
        iK[e] = np.linalg.solve(L.T, np.linalg.solve(L, eye)) - iB
    return iK, beta


def predict_given_factorizations(C, ell, sf2, m, s, iK, beta):
    """pilco/models/mgpr.py:91-149.  ``C`` are the centres (X, or Z for SMGPR: smgpr.py:47-48)."""
    E, D = ell.shape
    N = C.shape[0]
    s4 = np.broadcast_to(s[None, None], (E, E, D, D))                     # :98
    inp = np.broadcast_to((C - m)[None], (E, N, D))                        # :99
    iL = np.stack([np.diag(1.0 / ell[e]) for e in range(E)])               # :102
    iN = inp @ iL
    B = iL @ s4[0] @ iL + np.eye(D)
    t = np.linalg.solve(B.transpose(0, 2, 1), iN.transpose(0, 2, 1)).transpose(0, 2, 1)   # :108-110
    lb = np.exp(-(iN * t).sum(-1) / 2.0) * beta                            # :112
    tiL = t @ iL
    c = sf2 / np.sqrt(np.linalg.det(B))
    M = (lb.sum(-1) * c)[:, None]
    V = (tiL.transpose(0, 2, 1) @ lb[:, :, None])[..., 0] * c[:, None]

    Rm = s4 @ np.stack([[np.diag(1.0 / ell[i] ** 2 + 1.0 / ell[j] ** 2) for j in range(E)]
                        for i in range(E)]) + np.eye(D)                    # :121-124
    Xa = inp[None, :, :, :] / (ell ** 2)[:, None, None, :]                 # :127
    X2 = -inp[:, None, :, :] / (ell ** 2)[None, :, None, :]                # :128
    Q = np.linalg.solve(Rm, s4) / 2.0                                      # :129
    Xs = ((Xa @ Q) * Xa).sum(-1)
    X2s = ((X2 @ Q) * X2).sum(-1)
    maha = -2.0 * ((Xa @ Q) @ X2.transpose(0, 1, 3, 2)) + Xs[:, :, :, None] + X2s[:, :, None, :]
    k = np.log(sf2)[:, None] - (iN ** 2).sum(-1) / 2.0                     # :135-136
    L = np.exp(k[:, None, :, None] + k[None, :, None, :] + maha)           # :137
    S = (np.broadcast_to(beta[:, None, None, :], (E, E, 1, N)) @ L
         @ np.broadcast_to(beta[None, :, :, None], (E, E, N, 1)))[:, :, 0, 0]
    diagL = np.stack([L[e, e] for e in range(E)])                          # :143
    S = S - np.diag((iK * diagL).sum((1, 2)))
    S = S / np.sqrt(np.linalg.det(Rm))
    S = S + np.diag(sf2)
    S = S - M @ M.T
    return M.T, S, V.T


def predict_on_noisy_inputs(X, Y, ell, sf2, sn2, m, s):
    """pilco/models/mgpr.py:77-79."""
    iK, beta = calculate_factorizations(X, Y, ell, sf2, sn2)
    return predict_given_factorizations(X, ell, sf2, m, s, iK, beta)


def sparse_predict_on_noisy_inputs(X, Z, Y, ell, sf2, sn2, m, s):
    """SMGPR path: smgpr.py:24-48 then mgpr.py:91-149 centred on Z."""
    iK, beta = fitc_factorizations(X, Z, Y, ell, sf2, sn2)
    return predict_given_factorizations(Z, ell, sf2, m, s, iK, beta)


def squash_sin(m, s, max_action=None):
    """pilco/controllers.py:13-36."""
    k = m.shape[1]
    if max_action is None:
        max_action = np.ones((1, k))
    else:
        max_action = max_action * np.ones((1, k))
    ds = np.diag(s)
    M = max_action * np.exp(-ds / 2.0) * np.sin(m)
    lq = -(ds[:, None] + ds[None, :]) / 2.0
    q = np.exp(lq)
    S = (np.exp(lq + s) - q) * np.cos(m.T - m) - (np.exp(lq - s) - q) * np.cos(m.T + m)
    S = max_action * max_action.T * S / 2.0
    C = max_action * np.diag(np.exp(-ds / 2.0) * np.cos(m)[0])
    return M, S, C.reshape(k, k)


def linear_action(W, b, m, s, squash=True, max_action=None):
    """pilco/controllers.py:46-58."""
    M = m @ W.T + b
    S = W @ s @ W.T
    V = W.T
    if squash:
        M, S, V2 = squash_sin(M, S, max_action)
        V = V @ V2
    return M, S, V


def rbf_action(Xc, Yc, ell, m, s, squash=True, max_action=None, sf2=None, sn2=None):
    """pilco/controllers.py:108-121 (variance 1.0 fixed :91-93; noise 1e-4 fixed :77-78)."""
    U = Yc.shape[1]
    sf2 = np.ones(U) if sf2 is None else sf2
    sn2 = 1e-4 * np.ones(U) if sn2 is None else sn2
    iK, beta = calculate_factorizations(Xc, Yc, ell, sf2, sn2)
    M, S, V = predict_given_factorizations(Xc, ell, sf2, m, s, 0.0 * iK, beta)
    S = S - np.diag(sf2 - 1e-6)
    if squash:
        M, S, V2 = squash_sin(M, S, max_action)
        V = V @ V2
    return M, S, V


def exponential_reward(m, s, W, t):
    """pilco/rewards.py:19-51."""
    D = m.shape[1]
    SW = s @ W
    iSpW = np.linalg.solve((np.eye(D) + SW).T, W.T).T
    muR = np.exp(-(m - t) @ iSpW @ (m - t).T / 2.0) / np.sqrt(np.linalg.det(np.eye(D) + SW))
    i2SpW = np.linalg.solve((np.eye(D) + 2.0 * SW).T, W.T).T
    r2 = np.exp(-(m - t) @ i2SpW @ (m - t).T) / np.sqrt(np.linalg.det(np.eye(D) + 2.0 * SW))
    sR = r2 - muR @ muR
    return muR.reshape(1, 1), sR.reshape(1, 1)


def propagate(m_x, s_x, action_fn, dynamics_fn):
    """pilco/models/pilco.py:138-153.  action_fn(m,s)->(M,S,V); dynamics_fn(m,s)->(M,S,V)."""
    m_u, s_u, c_xu = action_fn(m_x, s_x)
    m = np.concatenate([m_x, m_u], axis=1)
    s1 = np.concatenate([s_x, s_x @ c_xu], axis=1)
    s2 = np.concatenate([(s_x @ c_xu).T, s_u], axis=1)
    s = np.concatenate([s1, s2], axis=0)
    M_dx, S_dx, C_dx = dynamics_fn(m, s)
    M_x = M_dx + m_x
    S_x = S_dx + s_x + s1 @ C_dx + C_dx.T @ s1.T
    return M_x, S_x


def predict(m_x, s_x, n, action_fn, dynamics_fn, reward_fn):
    """pilco/models/pilco.py:118-136 -- reward is accumulated at the pre-step state."""
    total = np.zeros((1, 1))
    for _ in range(n):
        r = reward_fn(m_x, s_x)[0]
        m_x, s_x = propagate(m_x, s_x, action_fn, dynamics_fn)
        total = total + r
    return m_x, s_x, total
