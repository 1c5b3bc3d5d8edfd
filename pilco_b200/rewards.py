"""Rewards (drop-in for pilco/rewards.py).  ``ExponentialReward`` moments run on the device
(``pilco_exp_reward``); ``LinearReward``/``CombinedRewards`` are linear maps of the state moments."""
import numpy as np

from . import engine, _lib
from .params import Parameter, host


class ExponentialReward:
    """exp(-(x-t)^T W (x-t)/2) under x~N(m,s) (rewards.py:7-51; reward.m)."""

    def __init__(self, state_dim, W=None, t=None):
        self.state_dim = state_dim
        self.W = Parameter(np.reshape(W, (state_dim, state_dim)) if W is not None else np.eye(state_dim), trainable=False)
        self.t = Parameter(np.reshape(t, (1, state_dim)) if t is not None else np.zeros((1, state_dim)), trainable=False)

    def compute_reward(self, m, s):
        k = self.state_dim
        m = np.asarray(m, dtype=np.float64).reshape(1, k)
        s = np.asarray(s, dtype=np.float64).reshape(1, k, k)
        mu, sR = engine.exp_reward(np.asarray(self.W), np.asarray(self.t).reshape(k), m, s, variance=True)
        return host(mu).reshape(1, 1), host(sR).reshape(1, 1)

    def terms(self, coef=1.0, channel=_lib.CHANNEL_ADD):
        return [dict(kind=_lib.REWARD_EXP, coef=float(coef), channel=channel, W=np.asarray(self.W),
                     t=np.asarray(self.t).reshape(self.state_dim))]


class LinearReward:
    """W^T x (rewards.py:53-61)."""

    def __init__(self, state_dim, W):
        self.state_dim = state_dim
        self.W = Parameter(np.reshape(W, (state_dim, 1)), trainable=False)

    def compute_reward(self, m, s):
        W = np.asarray(self.W)
        m = np.asarray(m, dtype=np.float64).reshape(1, self.state_dim)
        s = np.asarray(s, dtype=np.float64).reshape(self.state_dim, self.state_dim)
        return m @ W, W.T @ s @ W

    def terms(self, coef=1.0, channel=_lib.CHANNEL_ADD):
        return [dict(kind=_lib.REWARD_LINEAR, coef=float(coef), channel=channel,
                     W=np.asarray(self.W).reshape(self.state_dim), t=None)]


class CombinedRewards:
    """Weighted sum of rewards (rewards.py:64-81)."""

    def __init__(self, state_dim, rewards=[], coefs=None):
        self.state_dim = state_dim
        self.base_rewards = rewards
        self.coefs = Parameter(coefs if coefs is not None else np.ones(len(rewards)), trainable=False)

    def compute_reward(self, m, s):
        total_mean, total_cov = 0, 0
        for reward, coef in zip(self.base_rewards, np.asarray(self.coefs)):
            mu, cov = reward.compute_reward(m, s)
            total_mean = total_mean + coef * np.asarray(mu)
            total_cov = total_cov + coef ** 2 * np.asarray(cov)
        return total_mean, total_cov

    def terms(self, coef=1.0, channel=_lib.CHANNEL_ADD):
        out = []
        for reward, c in zip(self.base_rewards, np.asarray(self.coefs)):
            out.extend(reward.terms(coef * float(c), channel))
        return out
