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
