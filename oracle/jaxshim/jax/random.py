"""jax.random stand-in: keys are paths in the split tree; every draw is served by the installed tape, so the random
inputs are the same arrays the oracle and the CUDA path consume.  jax.random.choice is restated:
    p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key)); searchsorted(p_cuml, r)"""
from __future__ import annotations

from . import numpy as jnp
from ._core import Array


class Key:
    __slots__ = ("path",)

    def __init__(self, path=()):
        self.path = tuple(path)

    def __repr__(self):
        return f"Key{self.path}"


def PRNGKey(seed):
    return Key((int(seed),))


def split(key, num=2):
    return [Key(key.path + (i,)) for i in range(int(num))]


class Tape:
    """Override in the driver.  `key.path` identifies the call site (see tests/golden/make_golden.py)."""

    def uniform(self, key, shape, minval, maxval):
        raise NotImplementedError

    def dirichlet(self, key, alpha):
        raise NotImplementedError

    def gumbel(self, key, shape):
        raise NotImplementedError


_tape = Tape()


def install_tape(tape):
    global _tape
    _tape = tape


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return Array(_tape.uniform(key, tuple(shape), minval, maxval), dtype=jnp.float32)


def dirichlet(key, alpha):
    return Array(_tape.dirichlet(key, alpha), dtype=jnp.float32)


def gumbel(key, shape=(), dtype=None):
    return Array(_tape.gumbel(key, tuple(shape)), dtype=jnp.float32)


def choice(key, a, shape=(), replace=True, p=None):
    assert p is not None
    if not replace:
        # jax.random.choice without replacement (jax 0.4.35 _src/random.py): Gumbel top-k,
        #   g = -gumbel(key, (n,)) - log(p);  ind = argsort(g)[:n_draws]
        n = int(a)
        g = -gumbel(key, (n,)) - jnp.log(p)
        return jnp.argsort(g)[: int(shape[0])]
    assert shape == ()
    p_cuml = jnp.cumsum(p)
    r = p_cuml[-1] * (1 - uniform(key, ()))
    return jnp.searchsorted(p_cuml, r)
