"""Factorisation at scale (round-1 VERDICT item 8): time and fp64 rate of pilco_gp_factorize (Gram, Cholesky, triangular
inverse, iK = L^-T L^-1, beta) at N = 300, 1000, 2000 (E = 10, D = 12) and pilco_fitc_factorize at N = 2000, M = 200, plus one
pilco_fitc_nlml evaluation.  Run twice to compare: default (multi-CTA Cholesky + DMMA GEMM) vs
PILCO_GEMM_DFMA=1 PILCO_CHOL_SINGLE=1 (round-1 kernels).   python scripts/factorize_timing.py  -> one JSON line"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pilco_b200 import engine                    # noqa: E402
from util import make_gp_problem                 # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"env": {k: os.environ.get(k) for k in ("PILCO_GEMM_DFMA", "PILCO_CHOL_SINGLE")}}
E, D = 10, 12
for N in (300, 1000, 2000):
    X, Y, ell, sf2, sn2 = make_gp_problem(N, D, E, seed=N)
    gp = engine.gp_factorize(X, Y, ell, sf2, sn2)
    ms = timeit(lambda: engine.gp_refactorize(gp))
    flops = E * (N ** 3 / 3.0 + N ** 3 / 3.0 + 2.0 * N ** 3)        # Cholesky + triangular inverse + L^-T L^-1 (algorithmic)
    ref = np.linalg.inv(gp.sf2.cpu().numpy()[0] * np.exp(-0.5 * (((X[:, None, :] - X[None, :, :]) / ell[0]) ** 2).sum(-1)) + sn2[0] * np.eye(N))
    err = float(np.abs(gp.iK[0, :N, :N].cpu().numpy() - ref).max() / np.abs(ref).max())
    out["gp_factorize_N%d" % N] = {"ms": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12, "iK_rel_err_vs_numpy": err,
                                   "ok": int(gp.info.max().item()) == 0}
N, M = 2000, 200
X, Y, ell, sf2, sn2 = make_gp_problem(N, D, E, seed=7)
Z = np.random.RandomState(1).rand(M, D)
out["fitc_factorize_N2000_M200"] = {"ms": timeit(lambda: engine.fitc_factorize(X, Z, Y, ell, sf2, sn2))}
ev = engine.FitcNlml(X, Y, M, 1)
ZB = np.tile(Z, (1, E, 1, 1))
out["fitc_nlml_N2000_M200_E10"] = {"ms": timeit(lambda: ev(ZB, ell[None], sf2[None], sn2[None]))}
print(json.dumps(out))
