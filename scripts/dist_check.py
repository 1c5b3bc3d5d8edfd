"""Multi-GPU check (run under torchrun, NCCL): optimize_policy with sharded restarts must give every rank the
same winner, identical to what a single rank computes with all restarts batched."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pilco.models import PILCO                       # noqa: E402
from pilco.controllers import RbfController          # noqa: E402
from pilco.rewards import ExponentialReward          # noqa: E402


def build():
    np.random.seed(1)
    Ds, U = 3, 1
    X0 = np.random.rand(80, Ds + U)
    A = np.random.rand(Ds + U, Ds)
    Y0 = 0.1 * np.sin(X0).dot(A)
    ctrl = RbfController(Ds, U, 10, max_action=2.0)
    p = PILCO((X0, Y0), controller=ctrl, horizon=8, reward=ExponentialReward(Ds, t=np.array([0.5, 0.5, 0.5])),
              m_init=X0[0:1, :Ds], S_init=0.05 * np.eye(Ds))
    for mod in p.mgpr.models:
        mod.likelihood.variance.assign(1e-3)
        mod.kernel.lengthscales.assign(np.ones(Ds + U) * 2.0)
    return p


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    p = build()
    np.random.seed(7)
    best = p.optimize_policy(maxiter=10, restarts=6)
    flat = torch.as_tensor(p.controller.get_flat(), device="cuda")
    allf = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allf, flat)
    same = all(torch.equal(allf[0], t) for t in allf)
    # single-process reference on rank 0 (all 6 restarts in one batch): destroy the group view by calling policy_opt directly
    from pilco_b200 import policy_opt
    ok_single = True
    if rank == 0:
        q = build()
        np.random.seed(7)
        saved = policy_opt._dist
        policy_opt._dist = lambda: (None, 0, 1)
        q.optimize_policy(maxiter=10, restarts=6)
        policy_opt._dist = saved
        ok_single = bool(np.array_equal(q.controller.get_flat(), p.controller.get_flat()))
    if rank == 0:
        print("DIST_CHECK world=%d same_on_all_ranks=%s equals_single_rank=%s best_reward=%.6f" % (world, same, ok_single, best))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
