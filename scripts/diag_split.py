"""Diagnostics (not a bench line): where does the forward+backward step time go?  For the metric config at R
restarts: plain forward / taped forward only / taped forward + reverse sweep, each as ONE captured graph with
nsplit sub-batches on parallel streams.   python scripts/diag_split.py [R]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402
from pilco_b200 import engine, _lib              # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = bench.CONFIGS["metric"]
H, Ds, U = cfg["H"], cfg["Ds"], cfg["U"]
wl = bench.make_workload(cfg)
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
pol = bench.make_policies(cfg, np.arange(R))
ones, noise = np.ones((R, U)), 1e-4 * np.ones((R, U))
rew = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=wl["W"], t=wl["t"])]
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


def make(grad):
    def f(lo, hi):
        pg = engine.gp_factorize(pol["Xc"][lo:hi], pol["Yc"][lo:hi], pol["lc"][lo:hi], ones[lo:hi], noise[lo:hi], need_iK=False, mode=1)
        sp = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=np.ones(U), gp=pg)
        return engine.RolloutPlan(gp, sp, rew, wl["m0"], wl["S0"], H, R=hi - lo, grad=grad)
    return f


def timed(fn, K=5, W=3):
    for _ in range(W):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(K):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / K


for nsplit in (1, 2, 4, 8):
    a = engine.SplitRollout(make(False), R, nsplit=nsplit)
    b = engine.SplitRollout(make(True), R, nsplit=nsplit)
    c = engine.SplitRollout(make(True), R, nsplit=nsplit, backward=True)
    ta, tb, tc = timed(a.replay), timed(b.replay), timed(c.replay)
    print("R=%d nsplit=%d  plain fwd %.2f ms | taped fwd %.2f ms | taped fwd+bwd %.2f ms  (bwd alone %.2f)  -> %.0f / %.0f / %.0f steps/s" % (
        R, nsplit, ta, tb, tc, tc - tb, R * H / ta * 1e3, R * H / tb * 1e3, R * H / tc * 1e3), flush=True)
    del a, b, c
    torch.cuda.empty_cache()
