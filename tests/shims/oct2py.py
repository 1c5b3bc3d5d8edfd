"""Test-side stand-in for oct2py (Octave is not installable offline): ``Oct2Py().<fn>(...)`` EXECUTES the
reference's own MATLAB oracle file ``<fn>.m`` -- the unmodified copies under ``tests/reference_tests/Matlab Code``
-- with the MATLAB-subset interpreter ``oracle/mrun.py`` and mimics oct2py's return conventions (2-D arrays;
``nout`` selects how many outputs come back).  Set ``PILCO_SHIM_MATLAB_PORT=1`` to dispatch to the hand
transcription ``oracle.matlab_port`` instead (debugging only).
"""
import logging
import os

import numpy as np

from oracle import mrun as _mrun

_MDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reference_tests", "Matlab Code")
_USE_PORT = os.environ.get("PILCO_SHIM_MATLAB_PORT", "0") == "1"


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


def _arg(a):
    """Python value -> what the interpreter takes: dict structs (recursively), 2-D float arrays, floats."""
    if isinstance(a, dict):
        return {k: _arg(v) for k, v in a.items()}
    if isinstance(a, (bool, int, float, np.integer, np.floating)):
        return float(a)
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return float(a)
    if a.ndim == 1:                       # oct2py sends 1-D arrays as row vectors; empty -> 0x0
        return a.reshape(1, -1) if a.size else np.zeros((0, 0))
    return a


def _out(x):
    """oct2py hands 1x1 results back as Python floats, everything else as (at least 2-D) float arrays"""
    x = np.asarray(x, dtype=np.float64)
    return float(x) if x.size == 1 else x


class Oct2Py:
    """Only what the reference's tests use: ``addpath`` (ignored: the vendored directory is fixed) and calling
    ``.m`` functions by attribute with ``nout=``."""

    def __init__(self, logger=None, **kw):
        self.logger = logger

    def addpath(self, path):
        return None

    def __getattr__(self, name):
        if name.startswith("_") or not os.path.exists(os.path.join(_MDIR, name + ".m")):
            raise AttributeError(name)

        def call(*args, nout=1, **kw):
            if _USE_PORT:
                from oracle import matlab_port as _mp
                res = getattr(_mp, name)(*args)
                return res[:nout] if isinstance(res, tuple) and nout < len(res) else res
            res = _mrun.run(name, *[_arg(a) for a in args], nout=nout, mdir=_MDIR)
            return tuple(_out(r) for r in res) if nout > 1 else _out(res)
        return call
