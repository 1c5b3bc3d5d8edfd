"""Test-side stand-in for the parts of gpflow the reference's tests/examples touch."""
import numpy as np
from pilco_b200.params import Parameter, set_trainable   # noqa: F401


class config:
    @staticmethod
    def default_float():
        return np.float64

    @staticmethod
    def default_int():
        return np.int32


default_float = config.default_float
