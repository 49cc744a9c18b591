"""jax.numpy stand-in (see ../README.md): only what the reference's MCTS path calls."""
from __future__ import annotations

import numpy as np

from ._core import Array, _binary, _canon_dtype, _pathmath, result_dtype, unwrap

float32, int32, uint32, uint8, bool_ = np.float32, np.int32, np.uint32, np.uint8, np.bool_
number = np.number
ndarray = Array


def array(x, dtype=None):
    return Array(x, dtype=dtype)


asarray = array


def _default(dtype, fallback):
    return _canon_dtype(dtype if dtype is not None else fallback)


def zeros(shape, dtype=None):
    return Array(np.zeros(shape, dtype=_default(dtype, np.float32)))


def full(shape, fill_value, dtype=None):
    fv = unwrap(fill_value)
    return Array(np.full(shape, fv, dtype=_default(dtype, np.asarray(fv).dtype if not isinstance(fv, (int, float)) else
                                                   (np.int32 if isinstance(fv, int) else np.float32))))


def zeros_like(x):
    return Array(np.zeros_like(unwrap(x)))


def full_like(x, fill_value, dtype=None):
    xv = unwrap(x)
    return Array(np.full(xv.shape, unwrap(fill_value), dtype=_default(dtype, xv.dtype)))


def arange(n):
    return Array(np.arange(int(n), dtype=np.int32))


def where(c, a, b):
    rd = result_dtype(a, b)
    return Array(np.where(np.asarray(unwrap(c)).astype(bool), np.asarray(unwrap(a)).astype(rd), np.asarray(unwrap(b)).astype(rd)))


def minimum(a, b):
    return _binary(np.minimum, a, b)


def maximum(a, b):
    return _binary(np.maximum, a, b)


def min(x, axis=None):  # noqa: A001
    return Array(x).min(axis=axis)


def max(x, axis=None):  # noqa: A001
    return Array(x).max(axis=axis)


def sum(x, axis=None):  # noqa: A001
    return Array(x).sum(axis=axis)


def argmax(x, axis=None):
    return Array(x).argmax(axis=axis)


def cumsum(x):
    v = unwrap(x)
    if v.dtype.kind == "f":  # sequential left-to-right float32 accumulation
        out, acc = np.empty_like(v), np.float32(0)
        for i, e in enumerate(v):
            acc = np.float32(acc + e)
            out[i] = acc
        return Array(out)
    return Array(np.cumsum(v).astype(np.int32))


def greater(a, b):
    return Array(a) > b


def logical_and(a, b):
    return Array(np.logical_and(np.asarray(unwrap(a)).astype(bool), np.asarray(unwrap(b)).astype(bool)))


def sqrt(x):
    v = np.asarray(unwrap(x))
    return Array(np.sqrt(v.astype(np.float32)))


def log(x):
    v = np.asarray(unwrap(x), dtype=np.float32)
    out = np.asarray(_pathmath().tz_logf(v), dtype=np.float32)
    return Array(np.where(v == 0, np.float32(-np.inf), out).astype(np.float32))  # jnp.log(0) = -inf


def exp(x):
    return Array(_pathmath().tz_expf(np.asarray(unwrap(x), dtype=np.float32)))


def finfo(x):
    return np.finfo(x.dtype if isinstance(x, Array) else x)


def broadcast_to(x, shape):
    return Array(np.broadcast_to(unwrap(x), shape).copy())


def searchsorted(a, v):
    return Array(np.searchsorted(unwrap(a), unwrap(v), side="left").astype(np.int32))


# ---- the slice core/memory/replay_memory.py adds -----------------------------------------------------------------
def ones(shape, dtype=None):
    return Array(np.ones(shape, dtype=_default(dtype, np.float32)))


def empty_like(x):  # XLA materialises `empty` as zeros
    return Array(np.zeros_like(unwrap(x)))


def unravel_index(indices, shape):
    return tuple(Array(np.asarray(a).astype(np.int32)) for a in np.unravel_index(np.asarray(unwrap(indices)), shape))


def argsort(x):
    return Array(np.argsort(np.asarray(unwrap(x)), kind="stable").astype(np.int32))
