"""jax.tree_util stand-in: dataclasses (chex), dicts (sorted keys), lists, tuples, None; everything else is a leaf."""
import dataclasses


def tree_map(f, tree, *rest):
    if tree is None:
        return None
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        kw = {fld.name: tree_map(f, getattr(tree, fld.name), *[getattr(r, fld.name) for r in rest])
              for fld in dataclasses.fields(tree)}
        return type(tree)(**kw)
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *[r[k] for r in rest]) for k in sorted(tree)}
    if isinstance(tree, (list, tuple)):
        out = [tree_map(f, x, *[r[i] for r in rest]) for i, x in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out
