"""Small driver for ncu: a few dynamics moment-match calls at the metric shape (R restarts batched)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from pilco_b200 import engine                  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = bench.make_workload()
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
D = gp.D
m = np.tile(np.concatenate([wl["m0"], np.zeros(D - len(wl["m0"]))]), (R, 1))
s = np.tile(0.1 * np.eye(D), (R, 1, 1))
for _ in range(4):
    out = engine.mm_forward(gp, m, s)
torch.cuda.synchronize()
print("ok", float(out[0].sum()))
