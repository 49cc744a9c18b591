"""Array type with the JAX semantics the reference's MCTS path relies on (see ../README.md)."""
from __future__ import annotations

import numpy as np

_F32, _I32, _U32, _U8, _BOOL = np.dtype(np.float32), np.dtype(np.int32), np.dtype(np.uint32), np.dtype(np.uint8), np.dtype(np.bool_)


def _canon_dtype(dt) -> np.dtype:
    """x64 disabled: 64-bit types narrow to 32 bits."""
    dt = np.dtype(dt)
    if dt == np.float64 or dt == np.float16:
        return _F32
    if dt == np.int64:
        return _I32
    if dt == np.uint64:
        return _U32
    return dt


def _pathmath():
    from oracle import mcts_numpy as M  # this path's definitions of exp / log / pow / float sum

    return M


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        """Functional scatter: negative indices wrap, out-of-bounds updates are dropped (JAX default)."""
        out = self._arr._v.copy()
        val = np.asarray(unwrap(value)).astype(out.dtype, copy=False)
        idx = self._idx if isinstance(self._idx, tuple) else (self._idx,)
        idx = tuple(unwrap(i) for i in idx)
        if all(np.ndim(i) == 0 for i in idx):
            norm = []
            for ax, i in enumerate(idx):
                i, size = int(i), out.shape[ax]
                i = i + size if i < 0 else i
                if not 0 <= i < size:
                    return Array(out)  # dropped
                norm.append(i)
            out[tuple(norm)] = val
            return Array(out)
        assert len(idx) == 1 and np.ndim(idx[0]) == 1, "shim supports scalar tuples or one 1-D index array"
        i = np.asarray(idx[0]).astype(np.int64)
        size = out.shape[0]
        i = np.where(i < 0, i + size, i)
        ok = (i >= 0) & (i < size)
        val = np.broadcast_to(val, (len(i),) + out.shape[1:])
        out[i[ok]] = val[ok]  # duplicates: last write wins (XLA leaves the order unspecified; such slots are erased)
        return Array(out)


def unwrap(x):
    return x._v if isinstance(x, Array) else x


def _kind(x):
    """(dtype, weak) of an operand."""
    if isinstance(x, Array):
        return x._v.dtype, False
    if isinstance(x, (bool, np.bool_)):
        return _BOOL, not isinstance(x, np.bool_)
    if isinstance(x, int):
        return _I32, True
    if isinstance(x, float):
        return _F32, True
    a = np.asarray(x)
    return _canon_dtype(a.dtype), False


def _cat(dt):
    return 0 if dt == _BOOL else (2 if dt.kind == "f" else 1)


def result_dtype(a, b) -> np.dtype:
    (da, wa), (db, wb) = _kind(a), _kind(b)
    if wa and not wb:
        da, db, wa, wb = db, da, wb, wa
    if not wa and wb:  # strong array with a weak Python scalar: the array's dtype wins within a category
        if _cat(db) <= _cat(da):
            return da
        return _F32 if _cat(db) == 2 else _I32
    if _cat(da) != _cat(db):
        return da if _cat(da) > _cat(db) else db
    if da == db:
        return da
    if da.kind == "f":
        return _F32
    return _I32  # mixed integer widths / signedness


def _binary(op, a, b, out_dtype=None, compare=False):
    rd = result_dtype(a, b)
    x = np.asarray(unwrap(a)).astype(rd)
    y = np.asarray(unwrap(b)).astype(rd)
    with np.errstate(all="ignore"):
        r = op(x, y)
    if compare:
        return Array(r)
    return Array(np.asarray(r).astype(out_dtype or rd))


class Array:
    __slots__ = ("_v",)
    __array_priority__ = 1000

    def __init__(self, v, dtype=None):
        v = np.asarray(unwrap(v))
        dt = _canon_dtype(dtype if dtype is not None else v.dtype)
        self._v = v.astype(dt) if v.dtype != dt else v

    # --- introspection ---
    shape = property(lambda s: s._v.shape)
    dtype = property(lambda s: s._v.dtype)
    ndim = property(lambda s: s._v.ndim)
    size = property(lambda s: s._v.size)
    at = property(lambda s: _At(s))

    def __array__(self, dtype=None, copy=None):
        return self._v if dtype is None else self._v.astype(dtype)

    def __repr__(self):
        return f"ShimArray({self._v!r})"

    def __len__(self):
        return len(self._v)

    def __iter__(self):
        return (Array(x) for x in self._v)

    def __bool__(self):
        return bool(self._v)

    def __int__(self):
        return int(self._v)

    __index__ = __int__

    def __float__(self):
        return float(self._v)

    def item(self):
        return self._v.item()

    # --- indexing: negative wraps, out-of-bounds clamps (JAX gather default) ---
    def __getitem__(self, idx):
        tup = idx if isinstance(idx, tuple) else (idx,)
        norm = []
        for ax, i in enumerate(tup):
            i = unwrap(i)
            if isinstance(i, slice) or i is None or i is Ellipsis:
                norm.append(i)
                continue
            a = np.asarray(i)
            if a.dtype == np.bool_:
                norm.append(a)
                continue
            size = self._v.shape[ax]
            a = a.astype(np.int64)
            a = np.where(a < 0, a + size, a)
            norm.append(np.clip(a, 0, size - 1))
        norm = tuple(int(n) if isinstance(n, np.ndarray) and n.ndim == 0 else n for n in norm)
        return Array(self._v[norm])

    # --- shape / dtype ---
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self._v.reshape(shape))

    def astype(self, dt):
        return Array(self._v, dtype=dt)

    # --- reductions ---
    def sum(self, axis=None):
        if self._v.dtype.kind == "f":
            assert axis in (None, -1) and self._v.ndim <= 1
            return Array(_pathmath().canon_sum(self._v))
        dt = _I32 if self._v.dtype.kind in "bi" else self._v.dtype
        return Array(self._v.sum(axis=axis).astype(dt))

    def max(self, axis=None):
        return Array(self._v.max(axis=axis))

    def min(self, axis=None):
        return Array(self._v.min(axis=axis))

    def argmax(self, axis=None):
        return Array(np.argmax(self._v, axis=axis).astype(_I32))  # first maximum, like XLA

    # --- arithmetic ---
    def __add__(self, o): return _binary(np.add, self, o)
    def __radd__(self, o): return _binary(np.add, o, self)
    def __sub__(self, o): return _binary(np.subtract, self, o)
    def __rsub__(self, o): return _binary(np.subtract, o, self)
    def __mul__(self, o): return _binary(np.multiply, self, o)
    def __rmul__(self, o): return _binary(np.multiply, o, self)
    def __neg__(self): return Array(-self._v)
    def __mod__(self, o): return _binary(np.remainder, self, o)  # jnp.remainder: sign of the divisor, like NumPy

    def __truediv__(self, o):
        rd = result_dtype(self, o)
        rd = rd if rd.kind == "f" else _F32
        return Array(np.true_divide(np.asarray(unwrap(self)).astype(rd), np.asarray(unwrap(o)).astype(rd)).astype(rd))

    def __rtruediv__(self, o):
        return Array(o, dtype=result_dtype(self, o)).__truediv__(self)

    def __pow__(self, o):
        assert self._v.dtype == _F32 and isinstance(o, (float, int)), "shim: only float32 ** python scalar"
        return Array(_pathmath().tz_powf(self._v, np.float32(o)))

    def __eq__(self, o): return _binary(np.equal, self, o, compare=True)
    def __ne__(self, o): return _binary(np.not_equal, self, o, compare=True)
    def __lt__(self, o): return _binary(np.less, self, o, compare=True)
    def __le__(self, o): return _binary(np.less_equal, self, o, compare=True)
    def __gt__(self, o): return _binary(np.greater, self, o, compare=True)
    def __ge__(self, o): return _binary(np.greater_equal, self, o, compare=True)
    __hash__ = None

    def __and__(self, o): return _binary(np.bitwise_and, self, o)
    def __rand__(self, o): return _binary(np.bitwise_and, o, self)
    def __or__(self, o): return _binary(np.bitwise_or, self, o)
    def __ror__(self, o): return _binary(np.bitwise_or, o, self)

    def __invert__(self):
        return Array(~self._v)
