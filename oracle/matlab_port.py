"""Loop-form numpy (fp64) restatement of the reference's MATLAB oracle.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Every function follows one ``.m`` file under ``/root/reference/tests/Matlab Code/`` and keeps
its loop structure (clarity over speed).  MATLAB conventions are kept on the interface:
column vectors are ``[k, 1]`` arrays, ``hyp`` is ``[D+2, E]`` of *log* hyper-parameters
(log lengthscales, log signal std, log noise std).

Citations are ``file:line`` relative to ``/root/reference/tests/Matlab Code/``.
"""
import numpy as np


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


def maha(a, b, Q=None):
    """maha.m:22-29 -- pointwise squared Mahalanobis distance (a-b) Q (a-b)'."""
    if Q is None:
        return (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T
    aQ = a @ Q
    return (aQ * a).sum(1)[:, None] + ((b @ Q) * b).sum(1)[None, :] - 2.0 * aQ @ b.T


def _factorise_full(inputs, targets, hyp):
    """gp0.m:46-61 / gp2.m:50-67 -- K, inv(K+sn2 I), beta per output."""
    n, D = inputs.shape
    E = targets.shape[1]
    iK = np.zeros((n, n, E))
    beta = np.zeros((n, E))
    for i in range(E):
        inp = inputs / np.exp(hyp[:D, i])[None, :]
        K = np.exp(2.0 * hyp[D, i] - maha(inp, inp) / 2.0)
        L = np.linalg.cholesky(K + np.exp(2.0 * hyp[D + 1, i]) * np.eye(n))
        iK[:, :, i] = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(n)))
        beta[:, i] = np.linalg.solve(L.T, np.linalg.solve(L, targets[:, i]))
    return iK, beta


def _moments(centres, hyp, beta, iK, m, s, mode):
    """Shared body of gp0.m:63-104, gp1.m:85-124, gp2.m:69-106.

    mode 'gp0'/'gp1': subtract the trace term on the diagonal and add sf2 (gp0.m:93-100);
    mode 'gp2': deterministic GP, no trace term, add 1e-6 jitter (gp2.m:98-101).
    ``centres`` are the training inputs (gp0/gp2) or the pseudo-inputs (gp1).
    """
    n, D = centres.shape
    E = beta.shape[1]
    m = _col(m)
    k = np.zeros((n, E))
    M = np.zeros((E, 1))
    V = np.zeros((D, E))
    S = np.zeros((E, E))
    inp = centres - m.T                                           # gp0.m:65

    for i in range(E):                                            # gp0.m:68-81
        iL = np.diag(np.exp(-hyp[:D, i]))
        inn = inp @ iL
        B = iL @ s @ iL + np.eye(D)
        t = np.linalg.solve(B.T, inn.T).T                         # in/B
        l = np.exp(-(inn * t).sum(1) / 2.0)
        lb = l * beta[:, i]
        tiL = t @ iL
        c = np.exp(2.0 * hyp[D, i]) / np.sqrt(np.linalg.det(B))
        M[i, 0] = lb.sum() * c
        V[:, i] = tiL.T @ lb * c
        k[:, i] = 2.0 * hyp[D, i] - (inn * inn).sum(1) / 2.0

    for i in range(E):                                            # gp0.m:84-101
        ii = inp / np.exp(2.0 * hyp[:D, i])[None, :]
        for j in range(i + 1):
            R = s @ np.diag(np.exp(-2.0 * hyp[:D, i]) + np.exp(-2.0 * hyp[:D, j])) + np.eye(D)
            t = 1.0 / np.sqrt(np.linalg.det(R))
            ij = inp / np.exp(2.0 * hyp[:D, j])[None, :]
            L = np.exp(k[:, i][:, None] + k[:, j][None, :] + maha(ii, -ij, np.linalg.solve(R, s) / 2.0))
            if mode == 'gp2':
                S[i, j] = t * (beta[:, i] @ L @ beta[:, j])       # gp2.m:96
                S[j, i] = S[i, j]
            elif i == j:
                S[i, i] = t * (beta[:, i] @ L @ beta[:, i] - (iK[:, :, i] * L).sum())   # gp0.m:93
            else:
                S[i, j] = beta[:, i] @ L @ beta[:, j] * t         # gp0.m:95-96
                S[j, i] = S[i, j]
        if mode == 'gp2':
            S[i, i] += 1e-6                                       # gp2.m:99
        else:
            S[i, i] += np.exp(2.0 * hyp[D, i])                    # gp0.m:100

    S = S - M @ M.T                                               # gp0.m:104
    return M, S, V


def gp0(gpmodel, m, s):
    """gp0.m:36-104 -- exact multi-output GP moment match.  Returns M[E,1], S[E,E], V[D,E]."""
    hyp = np.asarray(gpmodel['hyp'], dtype=np.float64)
    inputs = np.asarray(gpmodel['inputs'], dtype=np.float64)
    targets = np.asarray(gpmodel['targets'], dtype=np.float64)
    iK, beta = _factorise_full(inputs, targets, hyp)
    return _moments(inputs, hyp, beta, iK, m, np.asarray(s, dtype=np.float64), 'gp0')


def gp2(gpmodel, m, s):
    """gp2.m:40-106 -- deterministic GP (RBF network) moment match."""
    hyp = np.asarray(gpmodel['hyp'], dtype=np.float64)
    inputs = np.asarray(gpmodel['inputs'], dtype=np.float64)
    targets = np.asarray(gpmodel['targets'], dtype=np.float64)
    iK, beta = _factorise_full(inputs, targets, hyp)
    return _moments(inputs, hyp, beta, iK, m, np.asarray(s, dtype=np.float64), 'gp2')


def gp1_factorise(gpmodel):
    """gp1.m:52-82 -- FITC factorisation: beta[np,E] and iK2[np,np,E] over pseudo-inputs."""
    ridge = 1e-6                                                  # gp1.m:42
    hyp = np.asarray(gpmodel['hyp'], dtype=np.float64)
    inputs = np.asarray(gpmodel['inputs'], dtype=np.float64)
    targets = np.asarray(gpmodel['targets'], dtype=np.float64)
    pinput = np.asarray(gpmodel['induce'], dtype=np.float64)
    if pinput.ndim == 2:
        pinput = pinput[:, :, None]
    n, D = inputs.shape
    E = targets.shape[1]
    npi, _, pE = pinput.shape
    beta = np.zeros((npi, E))
    iK2 = np.zeros((npi, npi, E))
    for i in range(E):
        pin = pinput[:, :, min(i, pE - 1)] / np.exp(hyp[:D, i])[None, :]
        inp = inputs / np.exp(hyp[:D, i])[None, :]
        Kmm = np.exp(2.0 * hyp[D, i] - maha(pin, pin) / 2.0) + ridge * np.eye(npi)
        Kmn = np.exp(2.0 * hyp[D, i] - maha(pin, inp) / 2.0)
        L = np.linalg.cholesky(Kmm)
        V = np.linalg.solve(L, Kmn)
        G = np.exp(2.0 * hyp[D, i]) - (V ** 2).sum(0)
        G = np.sqrt(1.0 + G / np.exp(2.0 * hyp[D + 1, i]))
        V = V / G[None, :]
        Am = np.linalg.cholesky(np.exp(2.0 * hyp[D + 1, i]) * np.eye(npi) + V @ V.T)
        At = L @ Am
        iAt = np.linalg.solve(At, np.eye(npi))
        iKi = (np.linalg.solve(Am, V / G[None, :]).T @ iAt).T     # gp1.m:77  [np, n]
        beta[:, i] = iKi @ targets[:, i]
        iB = iAt.T @ iAt * np.exp(2.0 * hyp[D + 1, i])
        iK2[:, :, i] = np.linalg.solve(Kmm, np.eye(npi)) - iB
    return iK2, beta, pinput


def gp1(gpmodel, m, s):
    """gp1.m:37-124 -- FITC sparse GP moment match (one shared pseudo-input set, pE=1)."""
    if 'induce' not in gpmodel or np.size(gpmodel['induce']) == 0:
        return gp0(gpmodel, m, s)                                 # gp1.m:38-39
    iK2, beta, pinput = gp1_factorise(gpmodel)
    assert pinput.shape[2] == 1, "per-output pseudo-inputs (pE>1) are not used by the reference tests"
    hyp = np.asarray(gpmodel['hyp'], dtype=np.float64)
    return _moments(pinput[:, :, 0], hyp, beta, iK2, m, np.asarray(s, dtype=np.float64), 'gp1')


def conlin(policy, m, s):
    """conlin.m:50-62 -- affine controller u = W x + b."""
    w = np.asarray(policy['p']['w'], dtype=np.float64)
    b = _col(policy['p']['b'])
    m = _col(m)
    M = w @ m + b
    S = w @ s @ w.T
    S = (S + S.T) / 2.0
    V = w.T.copy()
    return M, S, V


def gSin(m, v, e=None):
    """gSin.m:33-48 -- moments of e*sin(x), x ~ N(m, v), all indices squashed."""
    m = _col(m)
    v = np.asarray(v, dtype=np.float64)
    d = m.shape[0]
    if e is None:
        e = np.ones((d, 1))
    else:
        e = np.asarray(e, dtype=np.float64) * np.ones((d, 1))
    vii = np.diag(v).reshape(-1, 1)
    M = e * np.exp(-vii / 2.0) * np.sin(m)
    lq = -(vii + vii.T) / 2.0
    q = np.exp(lq)
    V = (np.exp(lq + v) - q) * np.cos(m - m.T) - (np.exp(lq - v) - q) * np.cos(m + m.T)
    V = (e @ e.T) * V / 2.0
    C = np.diag((e * np.exp(-vii / 2.0) * np.cos(m))[:, 0])
    return M, V, C


def reward(m, S, z, W):
    """reward.m:35-57 -- mean/variance (+analytic mean derivatives) of exp(-(x-z)'W(x-z)/2)."""
    m = _col(m)
    z = _col(z)
    D = m.shape[0]
    SW = S @ W
    iSpW = np.linalg.solve((np.eye(D) + SW).T, W.T).T             # W/(I+SW)
    muR = (np.exp(-(m - z).T @ iSpW @ (m - z) / 2.0) / np.sqrt(np.linalg.det(np.eye(D) + SW))).item()
    dmuRdm = -muR * (m - z).T @ iSpW                              # reward.m:48
    dmuRdS = muR * (iSpW @ (m - z) @ (m - z).T - np.eye(D)) @ iSpW / 2.0   # reward.m:49
    i2SpW = np.linalg.solve((np.eye(D) + 2.0 * SW).T, W.T).T
    r2 = (np.exp(-(m - z).T @ i2SpW @ (m - z)) / np.sqrt(np.linalg.det(np.eye(D) + 2.0 * SW))).item()
    sR = r2 - muR ** 2
    if sR < 1e-12:
        sR = 0.0                                                  # reward.m:56
    return muR, dmuRdm, dmuRdS, sR


def propagate(m, s, plant, dynmodel, policy, dyn_fn=gp0):
    """propagate.m:33-85 with angi=[] (no trig augmentation), noise disabled (:53-56),
    linear policy + sin squashing (:62-66), difference model (:82-85)."""
    m = _col(m)
    s = np.asarray(s, dtype=np.float64)
    poli = np.asarray(plant['poli'], dtype=int).ravel() - 1
    dyni = np.asarray(plant['dyni'], dtype=int).ravel() - 1
    difi = np.asarray(plant['difi'], dtype=int).ravel() - 1
    maxU = np.asarray(policy['maxU'], dtype=np.float64)
    D0 = m.shape[0]
    D1 = D0
    D2 = D1 + maxU.size
    D3 = D2 + D0
    M = np.zeros((D3, 1))
    M[:D0] = m
    S = np.zeros((D3, D3))
    S[:D0, :D0] = s

    i = poli
    j = np.arange(D1)
    k = np.arange(D1, D2)
    Mu, Su, C = conlin(policy, M[i], S[np.ix_(i, i)])             # propagate.m:62
    Mk, Sk, C2 = gSin(Mu, Su, maxU)                               # propagate.m:63
    M[k] = Mk
    S[np.ix_(k, k)] = Sk
    C = C @ C2
    q = S[np.ix_(j, i)] @ C
    S[np.ix_(j, k)] = q
    S[np.ix_(k, j)] = q.T

    ii = np.concatenate([dyni, np.arange(D1, D2)])                # propagate.m:69
    j = np.arange(D2)
    k = np.arange(D2, D3)
    Mk, Sk, C = dyn_fn(dynmodel, M[ii], S[np.ix_(ii, ii)])        # propagate.m:76
    M[k] = Mk
    S[np.ix_(k, k)] = Sk
    q = S[np.ix_(j, ii)] @ C
    S[np.ix_(j, k)] = q
    S[np.ix_(k, j)] = q.T

    P = np.hstack([np.zeros((D0, D2)), np.eye(D0)])               # propagate.m:84
    P[np.ix_(difi, difi)] = np.eye(difi.size)
    Mnext = P @ M
    Snext = P @ S @ P.T
    Snext = (Snext + Snext.T) / 2.0
    return Mnext, Snext


def pred(policy, plant, dynmodel, m, s, H, dyn_fn=gp0):
    """pred.m:29-39 -- H-step trajectory of marginals.  M[D,H+1], S[D,D,H+1]."""
    m = _col(m)
    D = m.shape[0]
    S = np.zeros((D, D, H + 1))
    M = np.zeros((D, H + 1))
    M[:, 0] = m[:, 0]
    S[:, :, 0] = s
    for i in range(H):
        m, s = propagate(m, s, plant, dynmodel, policy, dyn_fn)
        M[:, i + 1] = m[-D:, 0]
        S[:, :, i + 1] = s[-D:, -D:]
    return M, S
