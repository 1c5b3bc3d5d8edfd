"""PILCO orchestrator (drop-in for pilco/models/pilco.py:15-160).

The H-step moment-matching cascade (``predict``/``propagate``, pilco.py:118-153) and the policy
objective with its gradient run on the device through ``pilco_rollout_*``; policy optimisation batches
the random restarts (pilco.py:98-107) in one device rollout and shards them across ranks
(``pilco_b200.policy_opt``)."""
import time

import numpy as np
import pandas as pd

from .. import controllers, rewards, engine, _lib
from ..params import host, set_trainable
from .mgpr import MGPR
from .smgpr import SMGPR


class PILCO:
    def __init__(self, data, num_induced_points=None, horizon=30, controller=None,
                 reward=None, m_init=None, S_init=None, name=None):
        self.name = name
        if num_induced_points is None:
            self.mgpr = MGPR(data)
        else:
            self.mgpr = SMGPR(data, num_induced_points)
        self.state_dim = data[1].shape[1]
        self.control_dim = data[0].shape[1] - data[1].shape[1]
        self.horizon = horizon
        self.controller = controller if controller is not None else \
            controllers.LinearController(self.state_dim, self.control_dim)
        self.reward = reward if reward is not None else rewards.ExponentialReward(self.state_dim)
        if m_init is None or S_init is None:
            # first state of the data set, pilco.py:36-41
            self.m_init = data[0][0:1, 0:self.state_dim]
            self.S_init = np.diag(np.ones(self.state_dim) * 0.1)
        else:
            self.m_init = m_init
            self.S_init = S_init
        self.optimizer = None

    # ---- parameters ------------------------------------------------------------------------------
    @property
    def trainable_parameters(self):
        return list(self.mgpr.trainable_parameters) + list(self.controller.trainable_parameters)

    # ---- device rollout ---------------------------------------------------------------------------
    def policy_spec(self, flats=None):
        """Description of the policy for ``RolloutPlan``; ``flats`` [R,P] = per-restart flat parameters."""
        c = self.controller
        if isinstance(c, controllers.LinearController):
            return c.policy_spec(flats)
        if isinstance(c, controllers.RbfController):
            bf, Ds, U = c.policy_shapes
            if flats is None:
                gp = c.device_gp()
            else:
                X, Y, th = c.split_flat(flats)
                ell = c.models[0].kernel.lengthscales.transform.forward(th)      # lower + softplus (controllers.py:100)
                R = X.shape[0]
                # kernel variance and noise are the controller's own Parameters (sf2 = 1 and sn2 = 1e-4 unless the
                # user changed them, controllers.py:77-78,91-93) -- the same values compute_reward()/predict() see
                sf2 = np.tile(np.asarray(c.variance, dtype=np.float64).reshape(1, U), (R, 1))
                sn2 = np.tile(np.asarray(c.noise, dtype=np.float64).reshape(1, U), (R, 1))
                gp = engine.gp_factorize(X, Y, ell, sf2, sn2, need_iK=False, mode=1)
            return dict(kind=_lib.POLICY_RBF, Ds=Ds, U=U, squash=True,
                        max_action=controllers._max_action_vec(c.max_action, U), gp=gp)
        raise TypeError("unsupported controller type %r" % type(c))

    def reward_spec(self):
        """(reward terms, mult_mu) handed to the device rollout; SafePILCO adds a multiplicative channel."""
        return self.reward.terms(), 0.0

    def rollout_plan(self, m_x, s_x, n, flats=None):
        R = 1 if flats is None else int(np.asarray(flats).shape[0])
        terms, mult_mu = self.reward_spec()
        return engine.RolloutPlan(self.mgpr.device_gp(), self.policy_spec(flats), terms,
                                  np.asarray(m_x, dtype=np.float64).reshape(-1),
                                  np.asarray(s_x, dtype=np.float64), int(n), R=R, mult_mu=mult_mu)

    def _reward_key(self):
        terms, mult_mu = self.reward_spec()
        parts = [repr(float(mult_mu))]
        for t in terms:
            parts.append(repr((int(t["kind"]), float(t.get("coef", 1.0)), int(t.get("channel", 0)))))
            parts.append(np.ascontiguousarray(np.asarray(t["W"], dtype=np.float64)).tobytes())
            parts.append(b"" if t.get("t") is None else np.ascontiguousarray(np.asarray(t["t"], dtype=np.float64)).tobytes())
        return hash(tuple(parts))

    def _predict_plan(self, n):
        """Captured n-step cascade, cached per (model, policy, reward, n) state: the reference rebuilds nothing
        between calls either (tf.function), but here the cache is explicit -- any change of data, hyper-parameters,
        policy parameters or reward parameters changes the key and rebuilds the plan (a few ms)."""
        dyn = self.mgpr.device_gp()                       # (re)factorises only when data / hypers changed
        key = (id(dyn), self.mgpr._cache_key, self.controller.state_key(), self._reward_key(), int(n))
        cache = self.__dict__.setdefault("_plans", {})
        hit = cache.get(int(n))
        if hit is None or hit[0] != key:
            terms, mult_mu = self.reward_spec()
            plan = engine.PredictPlan(dyn, self.policy_spec(), terms, self.state_dim, int(n), mult_mu=mult_mu)
            if len(cache) >= 8:
                cache.clear()
            cache[int(n)] = hit = (key, plan)
        return hit[1]

    def predict(self, m_x, s_x, n):
        """n-step cascade (pilco.py:118-136) -> (m [1,Ds], S [Ds,Ds], reward [1,1]): one replay of a cached
        CUDA graph (host moments in, host moments out)."""
        from ..params import HostArray
        M, S, rew = self._predict_plan(n)(m_x, s_x)
        return M.view(HostArray), S.view(HostArray), rew.view(HostArray)

    def propagate(self, m_x, s_x):
        """one step (pilco.py:138-153) -> (M_x [1,Ds], S_x [Ds,Ds])"""
        M, S, _ = self.predict(m_x, s_x, 1)
        return M, S

    def training_loss(self):
        return -self.predict(self.m_init, self.S_init, self.horizon)[2]

    def compute_reward(self):
        return -self.training_loss()

    @property
    def maximum_log_likelihood_objective(self):
        return -self.training_loss()

    def compute_action(self, x_m):
        """pilco.py:115-116: the policy mean at a deterministic state (s = 0).  Served by a cached
        ``engine.ActionPlan`` (one captured CUDA graph per policy state) -- the per-control-step call of
        examples/utils.py:32-36."""
        from ..params import HostArray
        key = self.controller.state_key()
        if key != getattr(self, "_act_key", None):
            self._act_plan = engine.ActionPlan(self.policy_spec())
            self._act_key = key
        return self._act_plan(x_m).view(HostArray)

    # ---- model training (pilco.py:52-73) ------------------------------------------------------------
    def optimize_models(self, maxiter=200, restarts=1):
        self.mgpr.optimize(restarts=restarts)
        lengthscales, variances, noises = {}, {}, {}
        for i, model in enumerate(self.mgpr.models):
            lengthscales['GP' + str(i)] = model.kernel.lengthscales.numpy()
            variances['GP' + str(i)] = np.array([model.kernel.variance.numpy()])
            noises['GP' + str(i)] = np.array([model.likelihood.variance.numpy()])
        print('-----Learned models------')
        pd.set_option('display.precision', 3)
        print('---Lengthscales---')
        print(pd.DataFrame(data=lengthscales))
        print('---Variances---')
        print(pd.DataFrame(data=variances))
        print('---Noises---')
        print(pd.DataFrame(data=noises))

    # ---- policy optimisation (pilco.py:75-113) --------------------------------------------------------
    def optimize_policy(self, maxiter=50, restarts=1):
        from .. import policy_opt
        start = time.time()
        mgpr_trainable = self.mgpr.trainable_parameters
        for p in mgpr_trainable:
            set_trainable(p, False)
        try:
            best_flat, best_reward, all_rewards = policy_opt.optimize(self, maxiter=maxiter, restarts=restarts)
            self.controller.set_flat(best_flat)
        finally:
            for p in mgpr_trainable:
                set_trainable(p, True)
        end = time.time()
        for rwd in all_rewards:
            print("Controller's optimization: done in %.1f seconds with reward=%.3f." % (end - start, rwd))
        return best_reward
