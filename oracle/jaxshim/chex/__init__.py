"""chex stand-in: `dataclass` = frozen, keyword-only dataclasses.dataclass with .replace() (see ../README.md)."""
import dataclasses as _dc
from typing import Any

Array = Any
ArrayTree = Any
PRNGKey = Any
Numeric = Any


def dataclass(cls=None, *, frozen=False, **_):
    def wrap(c):
        c = _dc.dataclass(c, frozen=frozen, kw_only=True)
        c.replace = lambda self, **kw: _dc.replace(self, **kw)
        return c
    return wrap if cls is None else wrap(cls)
