"""Policies (drop-in for pilco/controllers.py): ``squash_sin``, ``LinearController``, ``RbfController``.
All moment computations run on the device through the C ABI (``pilco_squash_sin``,
``pilco_linear_action``, ``pilco_gp_factorize`` + ``pilco_mm_forward`` in deterministic-GP mode)."""
import numpy as np

from . import engine, _lib
from .models.mgpr import MGPR, Kernel
from .params import Parameter, Softplus, host, set_trainable


def _max_action_vec(max_action, k):
    if max_action is None:
        return np.ones(k)
    return np.broadcast_to(np.asarray(max_action, dtype=np.float64).reshape(-1), (k,)).copy()


def squash_sin(m, s, max_action=None):
    """Moments of max_action*sin(x), x~N(m,s) (controllers.py:13-36; gSin.m).  m [1,k], s [k,k]."""
    m = np.asarray(m, dtype=np.float64).reshape(1, -1)
    k = m.shape[1]
    s = np.asarray(s, dtype=np.float64).reshape(1, k, k)
    M, S, C = engine.squash_sin(m, s, _max_action_vec(max_action, k))
    return host(M), host(S[0]), host(C[0])


class LinearController:
    """u = W x + b (controllers.py:39-63)."""

    def __init__(self, state_dim, control_dim, max_action=1.0):
        self.W = Parameter(np.random.rand(control_dim, state_dim), name="W")
        self.b = Parameter(np.random.rand(1, control_dim), name="b")
        self.max_action = max_action

    @property
    def parameters(self):
        return [self.W, self.b]

    @property
    def trainable_parameters(self):
        return [p for p in self.parameters if p.trainable]

    def compute_action(self, m, s, squash=True):
        m = np.asarray(m, dtype=np.float64).reshape(1, -1)
        Ds = m.shape[1]
        s = np.asarray(s, dtype=np.float64).reshape(1, Ds, Ds)
        U = self.W.shape[0]
        M, S, V = engine.linear_action(np.asarray(self.W), np.asarray(self.b).reshape(U), m, s)
        if squash:
            M, S, C = engine.squash_sin(M, S, _max_action_vec(self.max_action, U))
            V = V @ C
        return host(M), host(S[0]), host(V[0])

    def randomize(self):
        mean, sigma = 0, 1
        self.W.assign(mean + sigma * np.random.normal(size=self.W.shape))
        self.b.assign(mean + sigma * np.random.normal(size=self.b.shape))

    # ---- flat parameter vector for the batched policy optimiser ---------------------------------
    def policy_spec(self, flat=None):
        U, Ds = self.W.shape
        if flat is None:
            W, b = np.asarray(self.W), np.asarray(self.b).reshape(U)
        else:
            flat = np.asarray(flat, dtype=np.float64)
            W = flat[..., :U * Ds].reshape(flat.shape[:-1] + (U, Ds))
            b = flat[..., U * Ds:].reshape(flat.shape[:-1] + (U,))
        return dict(kind=_lib.POLICY_LINEAR, Ds=Ds, U=U, squash=True, max_action=_max_action_vec(self.max_action, U),
                    W=W, b=b)

    def state_key(self):
        return hash((np.asarray(self.W).tobytes(), np.asarray(self.b).tobytes(),
                     np.asarray(self.max_action, dtype=np.float64).tobytes()))

    def get_flat(self):
        return np.concatenate([np.asarray(self.W).ravel(), np.asarray(self.b).ravel()])

    def set_flat(self, flat):
        U, Ds = self.W.shape
        self.W.assign(np.asarray(flat[:U * Ds]).reshape(U, Ds))
        self.b.assign(np.asarray(flat[U * Ds:]).reshape(1, U))


class FakeLikelihood:
    def __init__(self, variance):
        self.variance = Parameter(variance, transform=Softplus(1e-6), trainable=False, name="likelihood_variance")


class FakeGPR:
    """Parameter holder of one policy output (controllers.py:66-78): trainable centres X (shared) and
    targets Y, fixed noise 1e-4."""

    def __init__(self, data, kernel, X=None, likelihood_variance=1e-4):
        self.X = Parameter(data[0], name="DataX") if X is None else X
        self.Y = Parameter(data[1], name="DataY")
        self.data = [self.X, self.Y]
        self.kernel = kernel
        self.likelihood = FakeLikelihood(likelihood_variance)

    @property
    def parameters(self):
        return [self.X, self.Y, self.kernel.lengthscales, self.kernel.variance, self.likelihood.variance]

    @property
    def trainable_parameters(self):
        return [p for p in self.parameters if p.trainable]


class RbfController(MGPR):
    """RBF network policy = deterministic GP (controllers.py:80-129; gp2.m)."""
    mm_mode = 1

    def __init__(self, state_dim, control_dim, num_basis_functions, max_action=1.0):
        MGPR.__init__(self, [np.random.randn(num_basis_functions, state_dim),
                             0.1 * np.random.randn(num_basis_functions, control_dim)])
        for model in self.models:
            model.kernel.variance.assign(1.0)
            set_trainable(model.kernel.variance, False)
        self.max_action = max_action

    def create_models(self, data):
        self.models = []
        for i in range(self.num_outputs):
            kernel = Kernel(data[0].shape[1], lengthscale_lower=1e-3)        # controllers.py:100
            kernel.lengthscales.prior = (1.1, 1.0 / 10.0)
            if i == 0:
                self.models.append(FakeGPR((data[0], data[1][:, i:i + 1]), kernel))
            else:
                self.models.append(FakeGPR((data[0], data[1][:, i:i + 1]), kernel, self.models[-1].X))

    def compute_action(self, m, s, squash=True):
        """controllers.py:108-121 -> (M [1,U], S [U,U], V [Ds,U])"""
        gp = self.device_gp()
        m = np.asarray(m, dtype=np.float64).reshape(1, -1)
        s = np.asarray(s, dtype=np.float64).reshape(1, gp.D, gp.D)
        M, S, V, info = engine.mm_forward(gp, m, s)      # mode 1: no trace term, +1e-6 on the diagonal
        if squash:
            M, S, C = engine.squash_sin(M, S, _max_action_vec(self.max_action, gp.E))
            V = V @ C
        return host(M), host(S[0]), host(V[0])

    def randomize(self):
        print("Randomising controller")
        for m in self.models:
            m.X.assign(np.random.normal(size=m.data[0].shape))
            m.Y.assign(self.max_action / 10 * np.random.normal(size=m.data[1].shape))
            mean, sigma = 1, 0.1
            m.kernel.lengthscales.assign(mean + sigma * np.random.normal(size=m.kernel.lengthscales.shape))

    def state_key(self):
        return hash((self._state_key(), np.asarray(self.max_action, dtype=np.float64).tobytes()))

    # ---- flat parameter vector: [centres (bf*Ds) | targets (bf*U) | unconstrained lengthscales (U*Ds)] ----
    @property
    def policy_shapes(self):
        bf, Ds = self.models[0].X.shape
        return bf, Ds, self.num_outputs

    def get_flat(self):
        X = np.asarray(self.models[0].X).ravel()
        Y = self.Y.ravel()
        th = np.concatenate([m.kernel.lengthscales.unconstrained.ravel() for m in self.models])
        return np.concatenate([X, Y, th])

    def split_flat(self, flat):
        bf, Ds, U = self.policy_shapes
        flat = np.asarray(flat, dtype=np.float64)
        lead = flat.shape[:-1]
        X = flat[..., :bf * Ds].reshape(lead + (bf, Ds))
        Y = flat[..., bf * Ds:bf * Ds + bf * U].reshape(lead + (bf, U))
        th = flat[..., bf * Ds + bf * U:].reshape(lead + (U, Ds))
        return X, Y, th

    def set_flat(self, flat):
        X, Y, th = self.split_flat(flat)
        self.models[0].X.assign(X)
        for i, m in enumerate(self.models):
            m.Y.assign(Y[:, i:i + 1])
            m.kernel.lengthscales.set_unconstrained(th[i])
