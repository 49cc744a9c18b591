"""jax.lax control flow as plain Python (one un-batched env at a time)."""
from . import tree_util


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def scan(f, init, xs, length=None):
    carry, ys = init, []
    n = len(xs) if xs is not None else length
    for i in range(n):
        carry, y = f(carry, xs[i] if xs is not None else None)
        ys.append(y)
    return carry, (None if all(y is None for y in ys) else ys)
