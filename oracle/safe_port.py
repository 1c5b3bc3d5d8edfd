"""CPU restatement of the reference's safe-PILCO extension (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

numpy (scipy.stats.norm) and torch (autograd = gradient oracle) transcriptions of
  safe_pilco_extension/rewards_safe.py:13-73   RiskOfCollision, SingleConstraint, ObjectiveFunction
  safe_pilco_extension/safe_pilco.py:29-50     SafePILCO.predict (additive + multiplicative accumulators)

``tfd.Normal(loc, scale).cdf(x)`` is ``Phi((x - loc) / scale)``; the reference passes the covariance entry
(twice it for RiskOfCollision) as ``scale`` and that is reproduced as written.

PARITY UNPINNED: the reference holds no test, golden vector or MATLAB counterpart for this extension and
TensorFlow-Probability is not installable here, so this transcription is checked only against itself
(numpy vs torch, finite differences) -- see DESIGN.md section 8.
"""
import numpy as np
import torch
from scipy.stats import norm

F64 = torch.float64


# ---- numpy ---------------------------------------------------------------------------------------
def risk_of_collision(m, s, low, high):
    """rewards_safe.py:20-25"""
    infl = 2.0 * np.diag(s)
    c = lambda x, loc, scale: norm.cdf((x - loc) / scale)
    risk = (c(high[0], m[0, 0], infl[0]) - c(low[0], m[0, 0], infl[0])) * \
           (c(high[1], m[0, 2], infl[2]) - c(low[1], m[0, 2], infl[2]))
    return risk, 0.0001 * np.ones(1)


def single_constraint(m, s, dim, high=None, low=None, inside=True):
    """rewards_safe.py:28-58 (``high``/``low`` of None become False; the tests at :47,:50 are truthiness tests)."""
    high = False if high is None else high
    low = False if low is None else low
    c = lambda x: norm.cdf((x - m[0, dim]) / s[dim, dim])
    if not high:
        risk = 1.0 - c(low)
    elif not low:
        risk = c(high)
    else:
        risk = c(high) - c(low)
    if not inside:
        risk = 1.0 - risk
    return risk, 0.0001 * np.ones(1)


def objective_function(reward_fn, risk_fn, mu):
    """rewards_safe.py:60-73 -> callable (m, s) -> (reward - mu risk, var)"""
    def f(m, s):
        reward, var = reward_fn(m, s)
        risk, _ = risk_fn(m, s)
        return reward - mu * risk, var
    return f


def safe_predict(m_x, s_x, n, propagate_fn, reward_add_fn, reward_mult_fn, mu):
    """safe_pilco.py:29-50: the rewards are evaluated at the pre-step state of every step."""
    reward_add = np.zeros((1, 1))
    reward_mult = np.ones((1, 1))
    for _ in range(n):
        ra = reward_add_fn(m_x, s_x)[0]
        rm = reward_mult_fn(m_x, s_x)[0]
        m_x, s_x = propagate_fn(m_x, s_x)
        reward_add = reward_add + ra
        reward_mult = reward_mult * (1.0 - rm)
    return m_x, s_x, reward_add + mu * (1.0 - reward_mult)


# ---- torch (same formulas; autograd gives the gradient oracle) ---------------------------------
def _ndtr(x):
    return torch.special.ndtr(x)


def box_risk_torch(m, s, dims, lows, highs, sfac=1.0, inside=True):
    """General form of both risk classes: product over dims of Phi((high-m)/scale) - Phi((low-m)/scale)."""
    risk = torch.ones((), dtype=F64)
    for d, lo, hi in zip(dims, lows, highs):
        scale = sfac * s[d, d]
        ph = _ndtr((hi - m[0, d]) / scale) if np.isfinite(hi) else torch.ones((), dtype=F64)
        pl = _ndtr((lo - m[0, d]) / scale) if np.isfinite(lo) else torch.zeros((), dtype=F64)
        risk = risk * (ph - pl)
    return risk if inside else 1.0 - risk


def safe_predict_torch(m_x, s_x, n, propagate_fn, reward_add_fn, reward_mult_fn, mu):
    reward_add = torch.zeros((1, 1), dtype=F64)
    reward_mult = torch.ones((1, 1), dtype=F64)
    for _ in range(n):
        ra = reward_add_fn(m_x, s_x)
        rm = reward_mult_fn(m_x, s_x)
        m_x, s_x = propagate_fn(m_x, s_x)
        reward_add = reward_add + ra
        reward_mult = reward_mult * (1.0 - rm)
    return m_x, s_x, reward_add + mu * (1.0 - reward_mult)
