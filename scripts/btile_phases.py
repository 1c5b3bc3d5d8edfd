"""Phase breakdown of the BACKWARD tile kernel (diagnostics build: python -m pilco_b200.build --timing).
Per off-diagonal CTA: cycles in [entry -> row operands derived -> first TMA chunk landed -> column sweep done -> exit].
Usage (GPU box): PILCO_B200_LIB=pilco_b200/build_timing/libpilco_b200_timing.so python scripts/btile_phases.py [R]
(the DIAG launch that follows overwrites the first NB*E*R records; only later, off-diagonal records are read)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
from pilco_b200 import engine, _lib            # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
wl = bench.make_workload(bench.CONFIGS["metric"])
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
D, E = gp.D, gp.E
rng = np.random.RandomState(0)
m = np.tile(np.concatenate([wl["m0"], np.zeros(D - len(wl["m0"]))]), (R, 1))
s = np.tile(0.1 * np.eye(D), (R, 1, 1))
M, S, V, info = engine.mm_forward(gp, m, s)
gM, gS, gV = rng.randn(R, E), rng.randn(R, E, E), rng.randn(R, D, E)
for _ in range(2):
    engine.mm_backward(gp, m, s, M, gM, gS, gV)
torch.cuda.synchronize()
NB, P2 = 5, E * E
ncta = min(NB * P2 * R, 16384)
buf = (C.c_longlong * (5 * ncta))()
fn = _lib.lib.pilco_debug_btile_timing
fn.restype = C.c_int
rc = fn(buf, 5 * ncta)
t = np.frombuffer(buf, dtype=np.int64).reshape(ncta, 5)
idx = np.arange(ncta)
q = (idx // NB) % P2
keep = (idx >= NB * E * R) & (q // E != q % E)
d = np.diff(t[keep], axis=1).astype(np.float64)
res = {"rc": rc, "count": int(keep.sum()), "row_operands": float(d[:, 0].mean()), "tma_wait": float(d[:, 1].mean()),
       "sweep": float(d[:, 2].mean()), "epilogue": float(d[:, 3].mean()), "total": float(d.sum(1).mean())}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "btile_phases.json"), "w"), indent=1)
