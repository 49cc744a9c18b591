"""NumPy emulation of the slice of `jax` the reference's MCTS path uses -- see ../README.md.  NOT jax."""
from . import lax, nn, numpy, random, tree_util  # noqa: F401
from ._core import Array  # noqa: F401
from .tree_util import tree_map  # noqa: F401

__shim__ = True


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"jax.{name} is not emulated: the shim runs the reference one env at a time, eagerly")
    return f


jit, vmap, pmap = _unsupported("jit"), _unsupported("vmap"), _unsupported("pmap")
