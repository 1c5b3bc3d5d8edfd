"""Shared test helpers: seeded synthetic problems (SURVEY.md section 8d recipe) and error metrics."""
import numpy as np


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))


def scaled_err(a, b):
    """max |a-b| / max|b|  (robust when individual entries are ~0)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def make_gp_problem(n, D, E, seed=0, noise=1e-2):
    """X=rand(n,D), Y=sin(X)A+noise (test_predictions.py:19-21 recipe), benign fixed hyper-parameters."""
    rng = np.random.RandomState(seed)
    X = rng.rand(n, D)
    A = rng.rand(D, E)
    Y = np.sin(X).dot(A) + 1e-3 * (rng.rand(n, E) - 0.5)
    ell = 1.0 + rng.rand(E, D)
    sf2 = 1.0 + rng.rand(E)
    sn2 = noise * np.ones(E)
    return X, Y, ell, sf2, sn2


def make_input(D, seed=1, scale=1.0):
    rng = np.random.RandomState(seed)
    m = rng.rand(1, D)
    s = rng.rand(D, D)
    s = scale * s.dot(s.T)
    return m, s


def hyp_of(ell, sf2, sn2):
    """MATLAB log-hyper layout [D+2, E] (test_predictions.py:44-48)."""
    return np.log(np.hstack((ell, np.sqrt(sf2[:, None]), np.sqrt(sn2[:, None])))).T


def make_rollout_problem(N, Ds, U, bf, R, seed=0):
    """Seeded synthetic PILCO problem of a BASELINE.json config shape (SURVEY.md section 8d recipe, as bench.py builds the
    metric workload): dynamics data X=rand(N,D), state differences Y = 0.05 (sin(X) A + noise), fixed benign
    hyper-parameters; R RBF policies with bf centres.  Small state differences keep a 40-50 step rollout with
    untrained hyper-parameters finite (cf. tests/test_cascade.py:22)."""
    D = Ds + U
    rng = np.random.RandomState(seed)
    X = rng.rand(N, D)
    A = rng.rand(D, Ds)
    Y = 0.05 * (np.sin(X).dot(A) + 1e-3 * (rng.rand(N, Ds) - 0.5))
    ell = 1.0 + rng.rand(Ds, D)
    sf2 = 0.05 * (1.0 + rng.rand(Ds))
    sn2 = 1e-3 * np.ones(Ds)
    Xc, Yc, lc = [], [], []
    for r in range(R):
        rr = np.random.RandomState(seed + 1 + r)
        Xc.append(rr.randn(bf, Ds) * 0.5 + 0.5)
        Yc.append(0.1 * rr.randn(bf, U))
        lc.append(1.0 + 0.1 * rr.randn(U, Ds))
    return dict(X=X, Y=Y, ell=ell, sf2=sf2, sn2=sn2, Xc=np.stack(Xc), Yc=np.stack(Yc), lc=np.stack(lc),
                m0=X[0, :Ds].copy(), S0=0.1 * np.eye(Ds), W=np.eye(Ds), t=np.zeros(Ds), maxa=np.ones(U))


def oracle_rollout(P, r, H, dyn_fact=None, centres=None):
    """H-step cascade of restart r with the numpy port of the reference's Python path (oracle.python_port).
    ``dyn_fact`` = (iK, beta) to use instead of the exact-GP factorisation (FITC), with ``centres`` = Z."""
    from oracle import python_port as pp
    Ds = P["m0"].shape[0]
    iK, beta = dyn_fact if dyn_fact is not None else pp.calculate_factorizations(P["X"], P["Y"], P["ell"], P["sf2"], P["sn2"])
    C = P["X"] if centres is None else centres
    return pp.predict(P["m0"][None], P["S0"], H,
                      lambda m, s: pp.rbf_action(P["Xc"][r], P["Yc"][r], P["lc"][r], m, s, True, P["maxa"][None]),
                      lambda m, s: pp.predict_given_factorizations(C, P["ell"], P["sf2"], m, s, iK, beta),
                      lambda m, s: pp.exponential_reward(m, s, P["W"], P["t"][None]))
