"""End-to-end outer PILCO loop (examples/pendulum_numpy.py: data collection -> optimize_models on the device ->
batched optimize_policy -> acting with compute_action -> set_data) must run and improve the task reward."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pendulum_outer_loop():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import pendulum_numpy
    hist = pendulum_numpy.run(iters=2, T=20, restarts=3, maxiter=15, bf=8, verbose=False)
    pred = [h[0] for h in hist]
    ach = [h[1] for h in hist]
    assert np.all(np.isfinite(pred)) and np.all(np.isfinite(ach))
    # random torques hang around theta=0: reward about T*exp(-0.5*(pi/3)^2) = 11.6 at T=20; the learnt policy must beat it
    assert max(ach) > 12.5
