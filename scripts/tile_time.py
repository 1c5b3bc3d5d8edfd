import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from pilco_b200 import engine, _lib
from pilco_b200._lib import lib
from pilco_b200.engine import ptr, stream_ptr
R = 32
cfg = dict(bench.CONFIGS[os.environ.get("TT_CONFIG", "metric")])
if os.environ.get("TT_N"): cfg["N"] = int(os.environ["TT_N"])
wl = bench.make_workload(cfg)
gp = engine.gp_factorize(wl["X"], wl["Y"], wl["ell"], wl["sf2"], wl["sn2"])
N, D, E = gp.n, gp.D, gp.E; d = engine.device()
Mo = torch.empty((R, E), dtype=torch.float64, device=d); So = torch.empty((R, E, E), dtype=torch.float64, device=d)
Vo = torch.empty((R, D, E), dtype=torch.float64, device=d); info = torch.zeros(R, dtype=torch.int32, device=d)
wsb = lib.pilco_mm_workspace_bytes(N, D, E, R); ws = torch.empty(wsb // 8, dtype=torch.float64, device=d)
mj = engine.dev(np.tile(np.concatenate([wl["m0"], np.zeros(2)]), (R, 1))); sj = engine.dev(np.tile(0.1 * np.eye(D), (R, 1, 1)))
g = gp.struct(); ms3 = (C.c_float * 3)(); ts = []
for i in range(8):
    _lib.check(lib.pilco_mm_forward_profile(C.byref(g), R, ptr(mj), ptr(sj), ptr(Mo), ptr(So), ptr(Vo), ptr(info), ptr(ws), wsb, ms3, stream_ptr()))
    if i >= 3: ts.append(ms3[1])
print("N", cfg["N"], "setup_ms", round(float(ms3[0]), 4), "tile_ms", round(float(np.mean(ts)), 4))
