"""Generate tests/golden/configs.npz: expected results of full H-step rollouts at the BASELINE.json config shapes
(inverted_pendulum, inv_double_pendulum R=32, SMGPR N=2000/M=200, swimmer R=32 per GPU), computed
with the numpy port of the reference's Python path (oracle/python_port.py; pinned against the reference's MATLAB
oracle files by tests/test_golden.py).  The swimmer rollout alone costs ~90 s of CPU per restart, so the values are
committed instead of recomputed on the GPU box.  Problems come from tests/util.py:make_rollout_problem (seeded).

Run in the build container:   python tests/golden/make_config_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import python_port as pp                     # noqa: E402
from util import make_rollout_problem, oracle_rollout    # noqa: E402

# name -> (N, Ds, U, bf, H, R, seed, restarts checked)
CONFIGS = {
    "inverted_pendulum": (300, 4, 1, 10, 40, 1, 11, (0,)),
    "inv_double_pendulum": (400, 6, 1, 40, 40, 32, 12, (0, 31)),
    "swimmer": (500, 8, 2, 40, 50, 32, 13, (0, 31)),
    "metric": (300, 10, 2, 50, 40, 32, 14, (0, 31)),
}
SPARSE = dict(N=2000, M=200, Ds=10, U=2, bf=50, H=40, R=2, seed=15, check=(0, 1))


def main():
    out = {}
    for name, (N, Ds, U, bf, H, R, seed, check) in CONFIGS.items():
        P = make_rollout_problem(N, Ds, U, bf, R, seed=seed)
        for r in check:
            M, S, rew = oracle_rollout(P, r, H)
            out["%s_r%d_M" % (name, r)] = M
            out["%s_r%d_S" % (name, r)] = S
            out["%s_r%d_reward" % (name, r)] = rew
            print(name, r, float(rew[0, 0]), flush=True)
    # SMGPR (smgpr.py:24-52): FITC factorisation over M inducing points, then the same cascade over Z
    c = SPARSE
    P = make_rollout_problem(c["N"], c["Ds"], c["U"], c["bf"], c["R"], seed=c["seed"])
    Z = np.random.RandomState(c["seed"] + 100).rand(c["M"], c["Ds"] + c["U"])
    iK, beta = pp.fitc_factorizations(P["X"], Z, P["Y"], P["ell"], P["sf2"], P["sn2"])
    probe = np.random.RandomState(1).randn(c["M"], 3)
    out["sparse_beta"] = beta
    out["sparse_iK_probe"] = iK @ probe                   # [E,M,3]: a checksum of the [E,M,M] matrices
    out["sparse_iK_absmax"] = np.abs(iK).max()
    for r in c["check"]:
        M, S, rew = oracle_rollout(P, r, c["H"], dyn_fact=(iK, beta), centres=Z)
        out["sparse_r%d_M" % r] = M
        out["sparse_r%d_S" % r] = S
        out["sparse_r%d_reward" % r] = rew
        print("sparse", r, float(rew[0, 0]), flush=True)
    np.savez(os.path.join(HERE, "configs.npz"), **out)


if __name__ == "__main__":
    main()
