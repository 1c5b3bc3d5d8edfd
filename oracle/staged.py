"""numpy statement of the algorithm the CUDA kernels implement (forward stages + hand-derived VJPs).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The device kernels (pilco_b200/csrc/mm_kernels.cuh, mm_backward.cu, small_kernels.cuh, rollout_bwd.cu) are a
translation of the functions below; the tests check (1) these functions against ``python_port`` /
torch autograd on ``torch_port`` and (2) the device results against these functions, so a device bug can
be localised stage by stage.

Notation (DESIGN.md "Moment-match kernels"):  zeta_n = c_n - m,  p_a = ell_a^-2,  for the ORDERED pair (a,b):
  delta = p_a + p_b,  Amat = s + diag(1/delta),  Cm = Amat^-1,  Q = sym(0.5 diag(1/delta) Cm s),
  logdetR = logdet(Amat) + sum log delta,
  e[n,m] = k_a[n] + k_b[m] + za_n'Q za_n + zb_m'Q zb_m + 2 za_n'Q zb_m - 0.5 logdetR,   L' = exp(e),
  T_ab = sum_{n,m} (beta_a[n] beta_b[m] - d_ab iK_a[n,m]) L'[n,m].
"""
import numpy as np


def _sym(a):
    return 0.5 * (a + a.T)


def pair_setup(s, pa, pb):
    """Q, Cm, logdetR of one pair (device: mm_setup pair task)."""
    delta = pa + pb
    Amat = s + np.diag(1.0 / delta)
    Lc = np.linalg.cholesky(Amat)
    Cm = np.linalg.solve(Lc.T, np.linalg.solve(Lc, np.eye(len(pa))))
    Y = Cm @ s
    Q = _sym(0.5 * Y / delta[:, None])
    logdetR = 2.0 * np.log(np.diag(Lc)).sum() + np.log(delta).sum()
    return Q, Cm, logdetR


def mm_forward_staged(C, ell, sf2, beta, iK, m, s, mode=0):
    """Forward moment match in the kernels' formulation.  m [D], s [D,D] -> M [E], S [E,E], V [D,E]."""
    E, D = ell.shape
    s = _sym(s)
    zeta = C - m
    p = 1.0 / ell ** 2
    lsf2 = np.log(sf2)
    M = np.zeros(E)
    V = np.zeros((D, E))
    for a in range(E):                                       # output tasks (W-form)
        A = s + np.diag(ell[a] ** 2)
        W = np.linalg.inv(A)
        t = zeta @ W
        q = np.exp(-0.5 * (zeta * t).sum(1))
        c = np.exp(lsf2[a] + 0.5 * np.log(ell[a] ** 2).sum() - 0.5 * np.linalg.slogdet(A)[1])
        w = beta[a] * q
        M[a] = c * w.sum()
        V[:, a] = c * (w[:, None] * t).sum(0)
    k = lsf2[:, None] - 0.5 * (p[:, None, :] * zeta[None] ** 2).sum(-1)
    S = np.zeros((E, E))
    for b in range(E):
        for a in range(b + 1):
            Q, _, logdetR = pair_setup(s, p[a], p[b])
            za, zb = zeta * p[a], zeta * p[b]
            Ua = 2.0 * (za @ Q) * p[b]                               # U'[n] = p_b o (2 Q za_n)
            Ap = k[a] + ((za @ Q) * za).sum(1) - 0.5 * logdetR       # A'[n]
            Bq = k[b] + ((zb @ Q) * zb).sum(1)                       # B[m]
            Lp = np.exp(Ap[:, None] + Bq[None, :] + Ua @ zeta.T)
            G = np.outer(beta[a], beta[b])
            if a == b and mode == 0:
                G = G - iK[a]
            T = (G * Lp).sum()
            S[a, b] = S[b, a] = T
    S = S + np.diag(sf2 if mode == 0 else 1e-6 * np.ones(E)) - np.outer(M, M)
    return M, S, V


def mm_backward_staged(C, ell, sf2, beta, iK, m, s, gM, gS, gV, mode=0):
    """VJP of the moment match.  Returns gm [D], gs [D,D] (symmetric part), gC [n,D], gbeta [E,n], gell [E,D].

    Ordered-pair / row-side formulation: every ordered pair (a,b) is visited with weight
    gt_ab = gS[a,b] + gS[b,a] and only the derivatives through the ROW index n are taken; the column
    side of (a,b) is the row side of (b,a)."""
    E, D = ell.shape
    n = C.shape[0]
    s = _sym(s)
    zeta = C - m
    p = 1.0 / ell ** 2
    lsf2 = np.log(sf2)
    gzeta = np.zeros((n, D))
    gs = np.zeros((D, D))
    gbeta = np.zeros((E, n))
    gp = np.zeros((E, D))
    gell2 = np.zeros((E, D))

    # ---- forward recompute of M (needed by the centring term) -----------------------------------
    Wm, tm, qm, cm = [], [], [], []
    M = np.zeros(E)
    for a in range(E):
        A = s + np.diag(ell[a] ** 2)
        W = np.linalg.inv(A)
        t = zeta @ W
        q = np.exp(-0.5 * (zeta * t).sum(1))
        c = np.exp(lsf2[a] + 0.5 * np.log(ell[a] ** 2).sum() - 0.5 * np.linalg.slogdet(A)[1])
        Wm.append(W); tm.append(t); qm.append(q); cm.append(c)
        M[a] = c * (beta[a] * q).sum()
    gMtot = gM - (gS + gS.T) @ M

    # ---- mean / V block (device: mm_bfinish output tasks) -----------------------------------------
    for a in range(E):
        W, t, q, c = Wm[a], tm[a], qm[a], cm[a]
        w = beta[a] * q * c
        gw = gMtot[a] + t @ gV[:, a]
        glogc = (gw * w).sum()
        y = (w[:, None] * zeta).sum(0)
        gzeta += -(gw * w)[:, None] * t + w[:, None] * (W @ gV[:, a])[None, :]
        gW = -0.5 * (zeta * (gw * w)[:, None]).T @ zeta + _sym(np.outer(gV[:, a], y))
        gA = -W @ gW @ W - 0.5 * glogc * W
        gs += gA
        gbeta[a] += gw * q * c
        gell2[a] += np.diag(gA) + 0.5 * glogc / ell[a] ** 2

    # ---- covariance block: ordered pairs, row side (device: mm_btile + mm_bfinish pair tasks) ----
    k = lsf2[:, None] - 0.5 * (p[:, None, :] * zeta[None] ** 2).sum(-1)
    for a in range(E):
        for b in range(E):
            gt = gS[a, b] + gS[b, a]
            Q, Cm, logdetR = pair_setup(s, p[a], p[b])
            delta = p[a] + p[b]
            za, zb = zeta * p[a], zeta * p[b]
            Ua = 2.0 * (za @ Q) * p[b]
            Ap = k[a] + ((za @ Q) * za).sum(1) - 0.5 * logdetR
            Bq = k[b] + ((zb @ Q) * zb).sum(1)
            Lp = np.exp(Ap[:, None] + Bq[None, :] + Ua @ zeta.T)
            # tile pass outputs
            hL = Lp @ beta[b]                                   # sum_m beta_b[m] L'[n,m]
            HVL = (Lp * beta[b][None, :]) @ zeta                # sum_m beta_b[m] L'[n,m] zeta_m
            hr = beta[a] * hL
            HV = beta[a][:, None] * HVL
            if a == b and mode == 0:
                hr = hr - (iK[a] * Lp).sum(1)
                HV = HV - (iK[a] * Lp) @ zeta
            T = hr.sum()
            # finish
            gza = 2.0 * (hr[:, None] * za + HV * p[b][None, :]) @ Q          # sum_m H de/dza_n
            gzeta += gt * (p[a][None, :] * gza - hr[:, None] * za)
            Om = za.T @ (hr[:, None] * za) + za.T @ (HV * p[b][None, :])     # Omega^r_ab
            gQ = gt * _sym(Om)
            glogR = -0.25 * gt * T
            gs += 0.5 * Cm @ (gQ / delta[:, None] / delta[None, :]) @ Cm + glogR * Cm
            gbeta[a] += gt * hL
            gdelta = -2.0 * np.diag(Q @ gQ @ Q) + 2.0 * glogR * np.diag(Q)
            gp[a] += gt * ((gza * zeta).sum(0) - 0.5 * (hr[:, None] * zeta ** 2).sum(0)) + gdelta
            gp[b] += gdelta
    gell = -2.0 * gp / ell ** 3 + 2.0 * ell * gell2
    gm = -gzeta.sum(0)
    return gm, _sym(gs), gzeta, gbeta, gell


def mm_tape_forward(C, ell, sf2, beta, iK, m, s, mode=0):
    """What the TAPED forward tile pass leaves per unordered pair (a <= b) (device: mm_tape.cuh):
    Q, Cm, logdetR and, of H = G o L' (unweighted by any cotangent), the row sums hr, the column sums hc and the
    product H Z.  Returns a dict keyed by (a, b)."""
    E, D = ell.shape
    s = _sym(s)
    zeta = C - m
    p = 1.0 / ell ** 2
    lsf2 = np.log(sf2)
    k = lsf2[:, None] - 0.5 * (p[:, None, :] * zeta[None] ** 2).sum(-1)
    tape = {}
    for b in range(E):
        for a in range(b + 1):
            Q, Cm, logdetR = pair_setup(s, p[a], p[b])
            za, zb = zeta * p[a], zeta * p[b]
            Ua = 2.0 * (za @ Q) * p[b]
            Ap = k[a] + ((za @ Q) * za).sum(1) - 0.5 * logdetR
            Bq = k[b] + ((zb @ Q) * zb).sum(1)
            Lp = np.exp(Ap[:, None] + Bq[None, :] + Ua @ zeta.T)
            G = np.outer(beta[a], beta[b])
            if a == b and mode == 0:
                G = G - iK[a]
            H = G * Lp
            tape[(a, b)] = dict(Q=Q, Cm=Cm, logdetR=logdetR, hr=H.sum(1), hc=H.sum(0), HZ=H @ zeta)
    return tape


def mm_backward_tape(C, ell, sf2, beta, m, s, gM, gS, gV, tape):
    """VJP of the moment match w.r.t. the input moments only, from the tape (device: mm_tape_bfinish_kernel).
    Unordered pairs, both sides at once; no exponential of the N x N part is recomputed.
    Returns gm [D], gs [D,D] (symmetric part) -- equal to the first two results of mm_backward_staged."""
    E, D = ell.shape
    s = _sym(s)
    zeta = C - m
    p = 1.0 / ell ** 2
    lsf2 = np.log(sf2)
    gm = np.zeros(D)
    gs = np.zeros((D, D))
    # mean / V block: weighted moment sums with u_n = gw_n w_n, v_n = w_n
    Wm, M = [], np.zeros(E)
    for a in range(E):
        A = s + np.diag(ell[a] ** 2)
        W = np.linalg.inv(A)
        q = np.exp(-0.5 * ((zeta @ W) * zeta).sum(1))
        c = np.exp(lsf2[a] + 0.5 * np.log(ell[a] ** 2).sum() - 0.5 * np.linalg.slogdet(A)[1])
        Wm.append((W, beta[a] * q * c))
        M[a] = Wm[a][1].sum()
    gMtot = gM - (gS + gS.T) @ M
    for a in range(E):
        W, w = Wm[a]
        wgv = W @ gV[:, a]
        u = (gMtot[a] + zeta @ wgv) * w
        A1 = (zeta * u[:, None]).T @ zeta
        y1, y2 = zeta.T @ u, zeta.T @ w
        gW = -0.5 * A1 + _sym(np.outer(gV[:, a], y2))
        gs += -W @ gW @ W - 0.5 * u.sum() * W
        gm += W @ y1 - w.sum() * wgv
    # covariance block
    for b in range(E):
        for a in range(b + 1):
            t = tape[(a, b)]
            g = gS[a, a] if a == b else gS[a, b] + gS[b, a]
            Q, Cm, hr, hc, HZ = t["Q"], t["Cm"], t["hr"], t["hc"], t["HZ"]
            delta = p[a] + p[b]
            A1 = (zeta * hr[:, None]).T @ zeta
            A2 = (zeta * hc[:, None]).T @ zeta
            A3 = zeta.T @ HZ
            Pa, Pb = np.diag(p[a]), np.diag(p[b])
            X3 = Pa @ A3 @ Pb
            gQ = g * (Pa @ A1 @ Pa + Pb @ A2 @ Pb + X3 + X3.T)
            glogR = -0.5 * g * hr.sum()
            gs += 0.5 * Cm @ (gQ / delta[:, None] / delta[None, :]) @ Cm + glogR * Cm
            u = p[a] * (zeta.T @ hr) + p[b] * (zeta.T @ hc)
            gm += g * (u - 2.0 * delta * (Q @ u))
    return gm, _sym(gs)


# ---- closed forms -----------------------------------------------------------------------------------
def squash_forward(m, s, e):
    d = np.diag(s)
    M = e * np.exp(-d / 2.0) * np.sin(m)
    lq = -(d[:, None] + d[None, :]) / 2.0
    q = np.exp(lq)
    S = (np.exp(lq + s) - q) * np.cos(m[:, None] - m[None, :]) - (np.exp(lq - s) - q) * np.cos(m[:, None] + m[None, :])
    S = 0.5 * np.outer(e, e) * S
    Cd = e * np.exp(-d / 2.0) * np.cos(m)
    return M, S, np.diag(Cd)


def squash_backward(m, s, e, gM, gS, gC):
    """VJP of squash_sin (controllers.py:13-36); gC: only the diagonal matters."""
    U = len(m)
    M, S, C = squash_forward(m, s, e)
    Cd = np.diag(C)
    gm = gM * Cd - np.diag(gC) * M
    gs = np.zeros((U, U))
    gs[np.diag_indices(U)] += -0.5 * gM * M - 0.5 * np.diag(gC) * Cd
    d = np.diag(s)
    for i in range(U):
        for j in range(U):
            lq = -(d[i] + d[j]) / 2.0
            q = np.exp(lq)
            E1, E2 = np.exp(lq + s[i, j]), np.exp(lq - s[i, j])
            f = 0.5 * e[i] * e[j]
            cm_, cp_ = np.cos(m[i] - m[j]), np.cos(m[i] + m[j])
            sm_, sp_ = np.sin(m[i] - m[j]), np.sin(m[i] + m[j])
            g = gS[i, j]
            gs[i, j] += g * f * (E1 * cm_ + E2 * cp_)
            gs[i, i] += -0.5 * g * S[i, j]
            gs[j, j] += -0.5 * g * S[i, j]
            gm[i] += g * f * (-(E1 - q) * sm_ + (E2 - q) * sp_)
            gm[j] += g * f * ((E1 - q) * sm_ + (E2 - q) * sp_)
    return gm, gs


def exp_reward_grad(m, s, W, t):
    """muR and its derivatives (reward.m:48-49); W symmetric."""
    D = len(m)
    v = m - t
    A = np.eye(D) + s @ W
    Ainv = np.linalg.inv(A)
    iSpW = W @ Ainv
    wy = iSpW @ v
    mu = np.exp(-0.5 * v @ wy) / np.sqrt(np.linalg.det(A))
    dm = -mu * wy
    dS = 0.5 * mu * (np.outer(wy, wy) - iSpW)
    return mu, dm, _sym(dS)


def rbf_factor_backward(Xc, Yc, ell, gbeta, sn2=1e-4):
    """VJP through beta_a = (K_a + sn2 I)^-1 y_a (controllers.py:115 -> mgpr.py:81-89), sf2 = 1.
    Returns gX [bf,Ds], gY [bf,U], gell [U,Ds]."""
    bf, Ds = Xc.shape
    U = Yc.shape[1]
    gX = np.zeros((bf, Ds)); gY = np.zeros((bf, U)); gell = np.zeros((U, Ds))
    for a in range(U):
        diff = Xc[:, None, :] - Xc[None, :, :]
        K = np.exp(-0.5 * ((diff / ell[a]) ** 2).sum(-1))
        Kt = K + sn2 * np.eye(bf)
        beta = np.linalg.solve(Kt, Yc[:, a])
        gy = np.linalg.solve(Kt, gbeta[a])
        gY[:, a] = gy
        gK = -np.outer(gy, beta)
        P = (gK + gK.T) * K
        gX += -(P[:, :, None] * diff / ell[a] ** 2).sum(1)
        gell[a] = ((gK * K)[:, :, None] * diff ** 2).sum((0, 1)) / ell[a] ** 3
    return gX, gY, gell
