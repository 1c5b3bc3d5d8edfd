"""One forward (+ reverse) sweep of the metric-config rollout, eager, for ncu: `python scripts/profile_step.py
[--R 32] [--H 3] [--no-backward] [--config metric|swimmer|...]`.  Run twice inside (warm-up + profiled pass);
skip the first pass's launches with ncu -s."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pilco_b200 import engine, _lib            # noqa: E402
from util import make_rollout_problem           # noqa: E402

SHAPES = {"metric": (300, 10, 2, 50), "swimmer": (500, 8, 2, 40), "inv_double_pendulum": (400, 6, 1, 40),
          "inverted_pendulum": (300, 4, 1, 10)}
ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=32)
ap.add_argument("--H", type=int, default=3)
ap.add_argument("--config", default="metric")
ap.add_argument("--no-backward", dest="backward", action="store_false")
ap.add_argument("--passes", type=int, default=2)
a = ap.parse_args()
N, Ds, U, bf = SHAPES[a.config]
P = make_rollout_problem(N, Ds, U, bf, a.R, seed=0)
gp = engine.gp_factorize(P["X"], P["Y"], P["ell"], P["sf2"], P["sn2"])
pgp = engine.gp_factorize(P["Xc"], P["Yc"], P["lc"], np.ones((a.R, U)), 1e-4 * np.ones((a.R, U)), need_iK=False, mode=1)
spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=P["maxa"], gp=pgp)
plan = engine.RolloutPlan(gp, spec, [dict(kind=_lib.REWARD_EXP, coef=1.0, W=P["W"], t=P["t"])], P["m0"], P["S0"], a.H,
                          R=a.R, grad=a.backward)
torch.cuda.synchronize()
for _ in range(a.passes):
    plan.forward()
    if a.backward:
        plan.backward()
    torch.cuda.synchronize()
print("ok", float(plan.reward[0]))
