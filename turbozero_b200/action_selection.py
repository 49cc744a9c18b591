"""Action selectors -- mirrors core/evaluators/mcts/action_selection.py of the reference.

In the reference a selector is a Python callable traced into XLA, and its `q_transform` is any Python callable.  Here a
selector is a *descriptor* of a device functor compiled into the per-simulation kernel (csrc/tz_kernels.cu select_core /
narrow_select): it carries the kind and parameters that go into TzSearchCfg.  Arbitrary Python code cannot run inside the
kernel and there is no CPU fallback, so selectors and q_transforms come from two REGISTRIES of device functors:

  selectors      TZ_SEL_PUCT (PUCTSelector), TZ_SEL_MUZERO_PUCT (MuZeroPUCTSelector)          -- explore_scale / select_core
  q_transforms   TZ_QT_NORMALIZE (`normalize_q_values`, the default), TZ_QT_IDENTITY (`identity_q_values`)
                                                                                              -- q_transform_apply

Adding a q_transform (documented extension point):
  1. give it an id in include/tz_abi.h (TZ_QT_*, bump TZ_QT_COUNT) and a case in `q_transform_apply` (csrc/tz_kernels.cu);
  2. the same case in both oracles (oracle/mcts_numpy.py `q_transform`, oracle/tz_oracle.c `select_action`);
  3. `register_q_transform(host_fn, kind)` below, where `host_fn(q_values, child_n_values, parent_q_value, epsilon)` is the
     batched torch implementation users can call directly (the reference's functions are host-callable too).
Anything that is not registered raises at selector construction.
"""
from __future__ import annotations

from typing import Callable, Dict, Union

import torch

from ._abi import TZ_QT_IDENTITY, TZ_QT_NORMALIZE, TZ_SEL_MUZERO_PUCT, TZ_SEL_PUCT


def normalize_q_values(q_values: torch.Tensor, child_n_values: torch.Tensor, parent_q_value, epsilon: float) -> torch.Tensor:
    """action_selection.py:10-32, host-callable (torch, any device): min-max normalisation over ALL children and the
    parent, unvisited children completed with the minimum.  Shapes (..., F), (..., F), (...) or scalar.
    Inside a search the same arithmetic runs in the select kernel (TZ_QT_NORMALIZE)."""
    q_values = torch.as_tensor(q_values)
    parent = torch.as_tensor(parent_q_value, dtype=q_values.dtype, device=q_values.device)
    min_value = torch.minimum(parent, q_values.amin(dim=-1))
    max_value = torch.maximum(parent, q_values.amax(dim=-1))
    completed_by_min = torch.where(torch.as_tensor(child_n_values, device=q_values.device) > 0, q_values, min_value.unsqueeze(-1))
    denom = torch.clamp(max_value - min_value, min=epsilon)
    return (completed_by_min - min_value.unsqueeze(-1)) / denom.unsqueeze(-1)


def identity_q_values(q_values: torch.Tensor, child_n_values: torch.Tensor, parent_q_value, epsilon: float) -> torch.Tensor:
    """`lambda q, n, parent_q, eps: q`: the discounted child values as they are (TZ_QT_IDENTITY)."""
    return torch.as_tensor(q_values)


# host function (or its __name__) -> device functor id
_Q_TRANSFORMS: Dict[Union[str, Callable], int] = {}


def register_q_transform(host_fn: Callable, kind: int) -> Callable:
    """Declares that `host_fn` is implemented on the device as q_transform functor `kind` (include/tz_abi.h TZ_QT_*).
    Selectors accept the function object or its `__name__` as their `q_transform` argument."""
    _Q_TRANSFORMS[host_fn] = int(kind)
    _Q_TRANSFORMS[host_fn.__name__] = int(kind)
    return host_fn


register_q_transform(normalize_q_values, TZ_QT_NORMALIZE)
register_q_transform(identity_q_values, TZ_QT_IDENTITY)
_Q_TRANSFORMS["identity"] = TZ_QT_IDENTITY


def q_transform_kind(q_transform) -> int:
    try:
        return _Q_TRANSFORMS[q_transform]
    except (KeyError, TypeError):
        name = getattr(q_transform, "__name__", repr(q_transform))
        known = sorted(k for k in _Q_TRANSFORMS if isinstance(k, str))
        raise NotImplementedError(
            f"q_transform {name!r} has no device implementation (registered: {known}); arbitrary Python callables cannot run "
            "inside the select kernel and there is no host fallback -- see register_q_transform") from None


class MCTSActionSelector:
    """action_selection.py:35-58"""
    kind: int = -1

    def __init__(self, epsilon: float = 1e-8):
        self.epsilon = epsilon

    def __call__(self, tree, index, discount):
        raise NotImplementedError("selectors execute inside the select kernel (tz_select); see MCTS.traverse")

    def get_config(self) -> Dict:
        return {"epsilon": self.epsilon}

    def kernel_params(self) -> Dict:
        """{selector, c, c1, c2, epsilon, q_transform} for TzSearchCfg."""
        raise NotImplementedError(
            f"{type(self).__name__} has no device implementation: only PUCTSelector and MuZeroPUCTSelector are "
            "compiled into the sm_100a select kernel and there is no host fallback")


def _transform_name(q_transform) -> str:
    return q_transform if isinstance(q_transform, str) else q_transform.__name__


class PUCTSelector(MCTSActionSelector):
    """action_selection.py:61-116"""
    kind = TZ_SEL_PUCT

    def __init__(self, c: float = 1.0, epsilon: float = 1e-8, q_transform=normalize_q_values):
        super().__init__(epsilon=epsilon)
        self._qt = q_transform_kind(q_transform)
        self.c = c
        self.q_transform = q_transform

    def get_config(self) -> Dict:
        return {"c": self.c, 'q_transform': _transform_name(self.q_transform), **super().get_config()}

    def kernel_params(self) -> Dict:
        return dict(selector=self.kind, c=self.c, c1=0.0, c2=1.0, epsilon=self.epsilon, q_transform=self._qt)


class MuZeroPUCTSelector(MCTSActionSelector):
    """action_selection.py:119-177.  The reference's __call__ passes five arguments to a four-argument
    q_transform (:169) and so cannot run with its defaults; the kernel implements the intended maths
    u = p * sqrt(n) / (n_child + 1) * (log((n + c2 + 1) / c2) + c1) with the same normalised Q term.
    Pinned by tests/golden/muzero_puct_arityfix.npz: the reference's unmodified class run with a five-argument adapter
    around its own `normalize_q_values` passed as `q_transform`."""
    kind = TZ_SEL_MUZERO_PUCT

    def __init__(self, c1: float = 1.25, c2: float = 19652, epsilon: float = 1e-8, q_transform=normalize_q_values):
        super().__init__(epsilon=epsilon)
        self._qt = q_transform_kind(q_transform)
        self.c1 = c1
        self.c2 = c2
        self.q_transform = q_transform

    def get_config(self) -> Dict:
        return {"c1": self.c1, "c2": self.c2, "q_transform": _transform_name(self.q_transform), **super().get_config()}

    def kernel_params(self) -> Dict:
        return dict(selector=self.kind, c=1.0, c1=self.c1, c2=float(self.c2), epsilon=self.epsilon, q_transform=self._qt)
