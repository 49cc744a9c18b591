"""jax.nn.softmax = exp(x - max) / sum(exp(x - max)), with this path's exp and float-sum definitions."""
from ._core import Array, _pathmath, unwrap


def softmax(x, axis=-1):
    return Array(_pathmath().softmax(unwrap(x)))
