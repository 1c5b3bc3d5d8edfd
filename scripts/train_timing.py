"""GP hyper-parameter training: device (pilco_gp_nlml, all outputs x restarts in lock step) vs the host path
(torch-CPU autograd + SciPy per output) on the metric-size data set.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pilco.models import MGPR

N, D, E = 300, 12, 10
rng = np.random.RandomState(0)
X = rng.rand(N, D)
Y = np.sin(X).dot(rng.rand(D, E)) + 1e-2 * rng.randn(N, E)
out = {}
for name in ("device", "host"):
    np.random.seed(0)
    m = MGPR((X, Y))
    if name == "device":
        m.optimize(restarts=0, maxiter=1)        # warm-up (allocation, module load)
        m = MGPR((X, Y))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (m.optimize if name == "device" else m.optimize_host)(restarts=1)
    torch.cuda.synchronize()
    out[name + "_s"] = time.perf_counter() - t0
    out[name + "_loss"] = float(sum(mod.training_loss() for mod in m.models))
out["config"] = "N=%d D=%d E=%d, restarts=1 (2 initialisations per output), SciPy L-BFGS-B defaults" % (N, D, E)
print(json.dumps(out))
