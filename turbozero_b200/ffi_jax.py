"""jax.ffi registration of the C-ABI (north star: "a thin C-ABI registered with jax.ffi") -- the binding a maintainer of
the reference adds so that `core/evaluators/mcts/mcts.py` calls the sm_100a kernels instead of lowering to XLA.

NOT IMPORTED BY THE PACKAGE AND NOT TESTED IN THIS IMAGE: jax / jaxlib are not installed here (and cannot be: no network),
so this module raises ImportError on import.  Where jax >= 0.4.35 is available:

    python -c "from turbozero_b200.build import build_jax_ffi; build_jax_ffi()"   # csrc/tz_jax_ffi.cc -> lib/libtz_jax_ffi.so
    import turbozero_b200.ffi_jax as tzj                                          # registers the five FFI targets

Handlers and operand order: turbozero_b200/csrc/tz_jax_ffi.cc.  The derived tables of include/tz_abi.h (child_stats, best,
sel_state) travel as three extra leaves next to the reference's Tree leaves; `aux_init(N, F)` allocates them un-batched
(jax.vmap / init_batched add the env axis like for every other leaf).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - this image
    raise ImportError("turbozero_b200.ffi_jax needs jax (>= 0.4.35); the torch / ctypes binding in turbozero_b200 is the one "
                      "built and tested in this image") from e

if getattr(jax, "__shim__", False):  # oracle/jaxshim (test infrastructure) is not jax
    raise ImportError("turbozero_b200.ffi_jax needs the real jax; `jax` in this process is oracle/jaxshim")

try:  # jax >= 0.4.38
    from jax import ffi as _ffi
    _NEW_API = True
except ImportError:  # jax 0.4.35-0.4.37 (the reference pins 0.4.35): jax.extend.ffi, ffi_call(target, out, *args, **attrs)
    from jax.extend import ffi as _ffi
    _NEW_API = False

_LIB = Path(__file__).resolve().parent / "lib" / "libtz_jax_ffi.so"
_lib = ctypes.cdll.LoadLibrary(str(_LIB))
for _name in ("TzSetRoot", "TzSelect", "TzExpandBackprop", "TzRootAction", "TzReroot"):
    _ffi.register_ffi_target(_name, _ffi.pycapsule(getattr(_lib, _name)), platform="CUDA")

TZ_PATH_STRIDE = 66


def _call(target, out_types, args, aliases=None, **attrs):
    if _NEW_API:
        return _ffi.ffi_call(target, out_types, vmap_method="broadcast_all", input_output_aliases=aliases or {})(*args, **attrs)
    # jax 0.4.35-0.4.37: ffi_call has no aliasing argument, so the result buffers arrive uninitialised; the mutating
    # handlers copy every operand into its result when the two pointers differ (tz_jax_ffi.cc CopyIn)
    return _ffi.ffi_call(target, out_types, *args, vectorized=True, **attrs)


def aux_init(max_nodes: int, branching_factor: int):
    """The three derived tables for ONE tree (null rows {0, 0, 0, -1}; unknown decisions {-1, -1})."""
    cs = jnp.zeros((max_nodes, branching_factor, 4), jnp.int32).at[..., 3].set(-1)
    return cs, jnp.full((max_nodes, 2), -1, jnp.int32), jnp.zeros((8,), jnp.int32)


def tree_leaves(tree, aux, weighted=False):
    """Operand order of every handler (tz_jax_ffi.cc): Tree leaves, derived tables, optional r, embedding leaves."""
    d = tree.data
    emb = jax.tree_util.tree_leaves(d.embedding)
    head = [tree.next_free_idx, tree.parents, tree.edge_map, d.n, d.p, d.q, d.terminated, *aux]
    return head + ([d.r] if weighted else []) + emb, len(emb)


def _like(xs):
    return [jax.ShapeDtypeStruct(x.shape, x.dtype) for x in xs]


def select(tree, aux, path, selector_attrs, weighted=False):
    """MCTS.traverse (mcts.py:192-228) + parent-embedding gather (mcts.py:161-164).  Un-batched shapes; vmap adds B.
    Returns (parent, action, path, (best', sel_state'), parent embeddings): the walk fills in unknown best-table entries, so
    the caller threads the two returned derived tables back into `aux` (aux = (aux[0], best', sel_state'))."""
    leaves, k = tree_leaves(tree, aux, weighted)
    emb = leaves[-k:] if k else []
    outs = [jax.ShapeDtypeStruct((), jnp.int32), jax.ShapeDtypeStruct((), jnp.int32), jax.ShapeDtypeStruct(path.shape, path.dtype)]
    outs += _like([aux[1], aux[2]])
    outs += [jax.ShapeDtypeStruct(e.shape[1:], e.dtype) for e in emb]
    res = _call("TzSelect", outs, leaves + [path], aliases={len(leaves): 2, 8: 3, 9: 4}, weighted=int(weighted), n_emb=k,
                **selector_attrs)
    return res[0], res[1], res[2], (res[3], res[4]), list(res[5:])


def expand_backprop(tree, aux, parent, action, path, policy, value, terminated, new_emb, selector_attrs, weighted=False,
                    inv_q_temperature=1.0, fused=True, backprop_noise=None):
    """Second half of MCTS.iterate (mcts.py:174-189) [+ the next traverse when `fused`].  Returns the updated tree leaves (in
    handler order) and, when fused, (parent, action, path, parent embeddings) of the next simulation."""
    leaves, k = tree_leaves(tree, aux, weighted)
    new_leaves = jax.tree_util.tree_leaves(new_emb)
    args = leaves + [parent, action, path, policy, value, terminated] + new_leaves + ([backprop_noise] if backprop_noise is not None else [])
    outs = _like(leaves)
    aliases = {i: i for i in range(len(leaves))}
    if fused:
        outs += _like([parent, action, path]) + [jax.ShapeDtypeStruct(e.shape[1:], e.dtype) for e in leaves[len(leaves) - k:]]
        aliases.update({len(leaves) + j: len(leaves) + j for j in range(3)})
    res = _call("TzExpandBackprop", outs, args, aliases=aliases, weighted=int(weighted), inv_q_temperature=float(inv_q_temperature),
                n_emb=k, fused=int(fused), has_noise=int(backprop_noise is not None), **selector_attrs)
    return list(res[:len(leaves)]), list(res[len(leaves):])


def set_root(tree, aux, root_policy, root_value, root_embedding, weighted=False):
    """MCTS.update_root_node + Tree.set_root (mcts.py:363-384, tree.py:135-150)."""
    leaves, k = tree_leaves(tree, aux, weighted)
    args = leaves + [root_policy, root_value] + jax.tree_util.tree_leaves(root_embedding)
    return list(_call("TzSetRoot", _like(leaves), args, aliases={i: i for i in range(len(leaves))}, weighted=int(weighted), n_emb=k))


def root_action(tree, aux, noise, uniform01, temperature, weighted=False):
    """MCTS.sample_root_action + get_value (mcts.py:265-296, 111-120): (action, policy_weights, root q)."""
    leaves, k = tree_leaves(tree, aux, weighted)
    F = tree.edge_map.shape[-1]
    outs = [jax.ShapeDtypeStruct((), jnp.int32), jax.ShapeDtypeStruct((F,), jnp.float32), jax.ShapeDtypeStruct((), jnp.float32)]
    return _call("TzRootAction", outs, leaves + [noise, uniform01], temperature=float(temperature), weighted=int(weighted), n_emb=k)


def reroot(tree, aux, action, reset_flag, persist_tree=True, weighted=False):
    """MCTS.step / reset with the caller's reset-vs-step select (mcts.py:387-414, tree.py:169-278, common.py:89-94)."""
    leaves, k = tree_leaves(tree, aux, weighted)
    return list(_call("TzReroot", _like(leaves), leaves + [action, reset_flag], aliases={i: i for i in range(len(leaves))},
                      persist_tree=int(persist_tree), weighted=int(weighted), n_emb=k))
