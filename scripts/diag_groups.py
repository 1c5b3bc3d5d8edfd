"""Diagnostics (not a bench line): forward + reverse sweep with the restarts dealt to G independent GROUPS, each a
captured graph (nsplit sub-batches on parallel streams) replayed back to back on its own stream.  Restarts are
independent optimisation problems, so a group only has to wait for ITS OWN previous evaluation: with the groups half a
step out of phase the latency-bound reverse sweep of one group runs under the tile kernels of the other.
    python scripts/diag_groups.py [R] [K] [G:nsplit:stagger_ms,...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402
from pilco_b200 import engine, _lib              # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = bench.CONFIGS["metric"]
H, Ds, U = cfg["H"], cfg["Ds"], cfg["U"]
wl = bench.make_workload(cfg)
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
pol = bench.make_policies(cfg, np.arange(R))
ones, noise = np.ones((R, U)), 1e-4 * np.ones((R, U))
rew = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=wl["W"], t=wl["t"])]


def make(off):
    def f(lo, hi):
        lo, hi = lo + off, hi + off
        pg = engine.gp_factorize(pol["Xc"][lo:hi], pol["Yc"][lo:hi], pol["lc"][lo:hi], ones[lo:hi], noise[lo:hi], need_iK=False, mode=1)
        sp = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=np.ones(U), gp=pg)
        return engine.RolloutPlan(gp, sp, rew, wl["m0"], wl["S0"], H, R=hi - lo, grad=True)
    return f


def run(G, nsplit, stagger):
    per = R // G
    groups = [engine.SplitRollout(make(g * per), per, nsplit=nsplit, backward=True) for g in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    cur = torch.cuda.current_stream()
    clk = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000

    def bracket(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for g, st in enumerate(streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                if stagger and g:
                    torch.cuda._sleep(int(stagger * g * 1.9e6))       # stagger [ms] * g, in SM cycles
                for _ in range(k):
                    groups[g].graph.replay()
        for st in streams:
            cur.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    bracket(3)
    ms = bracket(K)
    print("R=%d groups=%d x nsplit=%d stagger=%.1f ms: %d steps in %.2f ms = %.2f ms/step -> %.0f steps/s" % (
        R, G, nsplit, stagger, K, ms, ms / K, R * H * K / ms * 1e3), flush=True)
    del groups
    torch.cuda.empty_cache()


CASES = ((1, 4, 0.0), (2, 2, 0.0), (2, 2, 7.0), (2, 4, 7.0), (2, 1, 7.0), (4, 1, 3.5), (4, 2, 3.5))
if len(sys.argv) > 3:                            # "G:nsplit:stagger_ms,..."
    CASES = tuple((int(a), int(b), float(c)) for a, b, c in (x.split(":") for x in sys.argv[3].split(",")))
for G, nsplit, stagger in CASES:
    if R % G == 0:
        run(G, nsplit, stagger)
