"""Driver for ncu: one forward + backward rollout at the metric shape (R restarts, H steps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pilco_b200 import engine, _lib
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.make_workload()
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
U, Ds = bench.CFG["U"], bench.CFG["Ds"]
Xc, Yc, lc = bench.make_policies(np.arange(R))
pgp = engine.gp_factorize(Xc, Yc, lc, np.ones((R, U)), 1e-4 * np.ones((R, U)), need_iK=False, mode=1)
spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=np.ones(U), gp=pgp)
plan = engine.RolloutPlan(gp, spec, [dict(kind=_lib.REWARD_EXP, coef=1.0, W=wl["W"], t=wl["t"])], wl["m0"], wl["S0"], H, R=R)
for _ in range(2):
    plan.forward(); plan.backward()
torch.cuda.synchronize()
print("ok", float(plan.reward.sum()))
