"""Phase breakdown of the forward tile kernel (diagnostics build: python -m pilco_b200.build --timing).
Per CTA: cycles in [entry -> row operands done -> first TMA chunk landed -> column sweep done -> exit].
Usage (GPU box): PILCO_B200_LIB=pilco_b200/build_timing/libpilco_b200_timing.so python scripts/tile_phases.py [R]"""
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
wl = bench.make_workload()
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
D = gp.D
m = np.tile(np.concatenate([wl["m0"], np.zeros(D - len(wl["m0"]))]), (R, 1))
s = np.tile(0.1 * np.eye(D), (R, 1, 1))
for _ in range(3):
    out = engine.mm_forward(gp, m, s)
torch.cuda.synchronize()
RPC = int(os.environ.get('PILCO_TILE_RPC', '5'))
NB, P = (5 + RPC - 1) // RPC, 55      # grid.x = row-block groups per pair
ncta = min(NB * P * R, 16384)
buf = (C.c_longlong * (5 * ncta))()
fn = _lib.lib.pilco_debug_tile_timing
fn.restype = C.c_int
rc = fn(buf, 5 * ncta)
t = np.frombuffer(buf, dtype=np.int64).reshape(ncta, 5)
d = np.diff(t, axis=1).astype(np.float64)
q = (np.arange(ncta) // NB) % P
b = np.floor((np.sqrt(8 * q + 1) - 1) / 2).astype(int)
a = q - b * (b + 1) // 2
diag = a == b
res = {"rc": rc, "ncta": int(ncta)}
for name, mask in (("offdiag", ~diag), ("diag", diag)):
    dd = d[mask]
    res[name] = {"count": int(mask.sum()),
                 "row_operands": float(dd[:, 0].mean()), "tma_wait": float(dd[:, 1].mean()),
                 "sweep": float(dd[:, 2].mean()), "epilogue": float(dd[:, 3].mean()),
                 "total": float(dd.sum(1).mean()), "sweep_p10": float(np.percentile(dd[:, 2], 10)),
                 "sweep_p90": float(np.percentile(dd[:, 2], 90))}
tot = d.sum(1)
res["cta_cycles_sum"] = float(tot.sum())
res["slot_cycles_per_cta_mean"] = float(tot.mean())
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tile_phases.json"), "w"), indent=1)
