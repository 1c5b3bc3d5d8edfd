"""Host-side parameter objects with the attribute surface user code expects from
``gpflow.Parameter`` / ``tf.Variable`` (``.numpy()``, ``.assign()``, ``.value()``, ``.shape``,
``.trainable``, ``.prior``, numpy interoperability), plus the softplus transforms GPflow applies to
positive parameters.  Pure host logic (numpy); no device code here.
"""
import numpy as np


class Softplus:
    """x = lower + log(1 + exp(theta))   (gpflow.utilities.positive(lower=...))"""

    def __init__(self, lower=0.0):
        self.lower = float(lower)

    def forward(self, theta):
        return self.lower + np.logaddexp(0.0, theta)

    def inverse(self, x):
        y = np.asarray(x, dtype=np.float64) - self.lower
        y = np.maximum(y, 1e-300)
        return y + np.log(-np.expm1(-y))

    def dforward(self, theta):
        return 1.0 / (1.0 + np.exp(-theta))


class Identity:
    lower = None

    def forward(self, theta):
        return theta

    def inverse(self, x):
        return np.asarray(x, dtype=np.float64)

    def dforward(self, theta):
        return np.ones_like(theta)


class Parameter:
    """A named fp64 array with an optional positivity transform, a trainable flag and a prior slot."""

    __array_priority__ = 100

    def __init__(self, value, transform=None, trainable=True, name=None, dtype=None, prior=None):
        if isinstance(value, Parameter):
            value = value.numpy()
        self._value = np.array(value, dtype=np.float64)
        self.transform = transform if transform is not None else Identity()
        self.trainable = bool(trainable)
        self.name = name
        self.prior = prior
        self.version = 0

    # --- tf.Variable-like surface -----------------------------------------------------------
    def numpy(self):
        return self._value.copy() if self._value.ndim else float(self._value)

    def value(self):
        return self._value.copy()

    def assign(self, value):
        v = np.asarray(value.numpy() if isinstance(value, Parameter) else value, dtype=np.float64)
        self._value = np.broadcast_to(v, self._value.shape).copy() if self._value.shape != v.shape else v.copy()
        self.version += 1
        return self

    @property
    def shape(self):
        return self._value.shape

    @property
    def unconstrained(self):
        return self.transform.inverse(self._value)

    def set_unconstrained(self, theta):
        self._value = np.asarray(self.transform.forward(np.asarray(theta, dtype=np.float64)), dtype=np.float64)
        self.version += 1

    def __array__(self, dtype=None, copy=None):
        return self._value.astype(dtype) if dtype is not None else self._value

    def __float__(self):
        return float(self._value)

    def __len__(self):
        return len(self._value)

    def __iter__(self):
        return iter(self._value)

    def __getitem__(self, idx):
        return self._value[idx]

    def __repr__(self):
        return "Parameter(%r, trainable=%s)" % (self._value, self.trainable)

    def _bin(self, other, op):
        o = other._value if isinstance(other, Parameter) else other
        return op(self._value, o)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, lambda a, b: np.add(b, a))
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, lambda a, b: np.subtract(b, a))
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, lambda a, b: np.multiply(b, a))
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: np.divide(b, a))
    def __neg__(self): return -self._value
    def __pow__(self, o): return self._bin(o, np.power)


def set_trainable(obj, flag):
    """gpflow.set_trainable: works on a Parameter or on any object exposing ``parameters``."""
    if isinstance(obj, Parameter):
        obj.trainable = bool(flag)
        return
    for p in getattr(obj, "parameters", []):
        p.trainable = bool(flag)


def positive(lower=0.0):
    return Softplus(lower)


class HostArray(np.ndarray):
    """ndarray returned at the API boundary; ``.numpy()`` mirrors TF eager tensors."""

    def numpy(self):
        return np.asarray(self)


def host(t):
    """CUDA tensor -> HostArray (device->host copy; synchronises the current stream)."""
    return t.detach().cpu().numpy().view(HostArray)
