"""Experiment: R restarts as `nsplit` sub-batches on separate streams inside ONE captured graph, so that the
latency-bound per-step kernels of one sub-batch overlap the tile kernel of another."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pilco_b200 import engine, _lib

R, H = 32, bench.CFG["H"]
wl = bench.make_workload()
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
U, Ds = bench.CFG["U"], bench.CFG["Ds"]
rew = [dict(kind=_lib.REWARD_EXP, coef=1.0, W=wl["W"], t=wl["t"])]
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for nsplit in (1, 2, 4):
    plans = []
    per = R // nsplit
    for k in range(nsplit):
        Xc, Yc, lc = bench.make_policies(np.arange(k * per, (k + 1) * per))
        pgp = engine.gp_factorize(Xc, Yc, lc, np.ones((per, U)), 1e-4 * np.ones((per, U)), need_iK=False, mode=1)
        spec = dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True, max_action=np.ones(U), gp=pgp)
        plans.append(engine.RolloutPlan(gp, spec, rew, wl["m0"], wl["S0"], H, R=per))
    for p in plans:
        p.forward()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(nsplit - 1)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        plans[0].forward()
        for s, p in zip(streams, plans[1:]):
            with torch.cuda.stream(s):
                p.forward()
        for s in streams:
            cur.wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    tot = 0.0
    K = 10
    for _ in range(K):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / K
    print("nsplit", nsplit, "ms/rollout", round(ms, 3), "steps/s", round(R * H / (ms * 1e-3)))
