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


def test_safe_cars_outer_loop():
    """examples/safe_cars_numpy.py (the reference's safe_cars_run.py scenario without gym): SafePILCO end to end --
    model training, policy optimisation with the multiplicative risk channel, prefix predictions, acting through
    compute_action, mu adaptation -- must run and produce finite, consistent numbers."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import safe_cars_numpy
    hist = safe_cars_numpy.run(iters=2, T=12, J=2, restarts=2, maxiter=8, bf=10, verbose=False)
    assert len(hist) == 2
    for h in hist:
        assert np.isfinite(h["objective"]) and np.isfinite(h["progress"])
        assert 0.0 <= h["overall_risk"] <= 1.0
    assert hist[1]["mu"] != hist[0]["mu"] or 0.025 <= hist[0]["overall_risk"] < 0.10     # mu adapts outside the dead band


@pytest.mark.parametrize("kind", ["linear", "rbf"])
def test_compute_action_fast_path(kind):
    """PILCO.compute_action (captured ActionPlan, s = 0) == controller.compute_action(x, 0)[0]
    (pilco/models/pilco.py:115-116), also after the policy parameters change."""
    from pilco.models import PILCO
    from pilco.controllers import RbfController, LinearController
    np.random.seed(3)
    Ds, U = 3, 2
    X0 = np.random.rand(40, Ds + U)
    Y0 = 0.1 * np.sin(X0).dot(np.random.rand(Ds + U, Ds))
    ctrl = LinearController(Ds, U, max_action=np.array([1.5, 0.5])) if kind == "linear" \
        else RbfController(Ds, U, 9, max_action=0.8)
    pilco = PILCO((X0, Y0), controller=ctrl, horizon=3)
    for trial in range(2):
        for _ in range(3):
            x = np.random.randn(1, Ds)
            u = pilco.compute_action(x)
            ref = ctrl.compute_action(x, np.zeros((Ds, Ds)))[0]
            assert u.shape == (1, U) and np.max(np.abs(np.asarray(u) - np.asarray(ref))) < 1e-12
            # ... and against the CPU oracle (numpy port of controllers.py:46-58 / 108-121), not only device vs device
            from oracle import python_port as pp
            maxa = np.broadcast_to(np.asarray(ctrl.max_action, dtype=np.float64).reshape(-1), (U,))[None]
            if kind == "linear":
                orc = pp.linear_action(np.asarray(ctrl.W), np.asarray(ctrl.b), x, np.zeros((Ds, Ds)), True, maxa)[0]
            else:
                orc = pp.rbf_action(np.asarray(ctrl.models[0].X), ctrl.Y, ctrl.lengthscales, x, np.zeros((Ds, Ds)), True, maxa)[0]
            assert np.max(np.abs(np.asarray(u) - np.asarray(orc))) < 1e-9
        ctrl.randomize()                      # new parameters -> the cached plan must be rebuilt
