"""Test-side stand-in for oct2py: ``Oct2Py()`` dispatches the MATLAB oracle calls of the reference's tests
to the numpy transcription ``oracle.matlab_port`` and mimics Octave's return conventions
(column vectors stay 2-D, 1x1 results become Python floats)."""
import logging

import numpy as np

from oracle import matlab_port as _mp


def get_log(name=None):
    return logging.getLogger(name or "oct2py")


class _Struct(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class io:
    Struct = _Struct


def _squeeze(x):
    x = np.asarray(x)
    return float(x) if x.size == 1 else x


class Oct2Py:
    def __init__(self, logger=None, **kw):
        self.logger = logger

    def addpath(self, path):
        return None

    def gp0(self, gpmodel, m, s, nout=3, **kw):
        return _mp.gp0(gpmodel, m, s)

    def gp1(self, gpmodel, m, s, nout=3, **kw):
        return _mp.gp1(gpmodel, m, s)

    def gp2(self, gpmodel, m, s, nout=3, **kw):
        return _mp.gp2(gpmodel, m, s)

    def conlin(self, policy, m, s, nout=3, **kw):
        return _mp.conlin(policy, m, s)

    def gSin(self, m, s, e, nout=3, **kw):
        return _mp.gSin(m, s, e)

    def reward(self, m, s, z, W, nout=4, **kw):
        return _mp.reward(m, s, z, W)[:nout]

    def pred(self, policy, plant, dynmodel, m, s, H, nout=2, **kw):
        return _mp.pred(policy, plant, dynmodel, m, s, int(H))
