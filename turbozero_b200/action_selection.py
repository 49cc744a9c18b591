"""Action selectors -- mirrors core/evaluators/mcts/action_selection.py of the reference.

In the reference a selector is a Python callable traced into XLA.  Here a selector is a *descriptor* of a device
functor compiled into the select kernel (csrc/tz_kernels.cu select_level): it carries the kind and parameters
that go into TzSearchCfg.  Arbitrary Python selectors cannot run inside the kernel and there is no CPU
fallback, so anything that is not one of the registered kinds raises at MCTS construction.
"""
from __future__ import annotations

from typing import Dict

from ._abi import TZ_SEL_MUZERO_PUCT, TZ_SEL_PUCT


def normalize_q_values(q_values, child_n_values, parent_q_value, epsilon):
    """action_selection.py:10-32.  Marker for the only q_transform the kernels implement (min-max
    normalisation over all F children and the parent, unvisited children completed with the minimum).
    It is evaluated on the device; calling it on the host is not part of the path."""
    raise NotImplementedError("normalize_q_values runs inside the sm_100a select kernel; it is not a host function")


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
        """{selector, c, c1, c2, epsilon} for TzSearchCfg."""
        raise NotImplementedError(
            f"{type(self).__name__} has no device implementation: only PUCTSelector and MuZeroPUCTSelector are "
            "compiled into the sm_100a select kernel and there is no host fallback")


def _check_transform(q_transform):
    if q_transform is not normalize_q_values:
        raise NotImplementedError("only q_transform=normalize_q_values is implemented by the select kernel")


class PUCTSelector(MCTSActionSelector):
    """action_selection.py:61-116"""
    kind = TZ_SEL_PUCT

    def __init__(self, c: float = 1.0, epsilon: float = 1e-8, q_transform=normalize_q_values):
        super().__init__(epsilon=epsilon)
        _check_transform(q_transform)
        self.c = c
        self.q_transform = q_transform

    def get_config(self) -> Dict:
        return {"c": self.c, 'q_transform': self.q_transform.__name__, **super().get_config()}

    def kernel_params(self) -> Dict:
        return dict(selector=self.kind, c=self.c, c1=0.0, c2=1.0, epsilon=self.epsilon)


class MuZeroPUCTSelector(MCTSActionSelector):
    """action_selection.py:119-177.  The reference's __call__ passes five arguments to a four-argument
    q_transform (:169) and so cannot run with its defaults; the kernel implements the intended maths
    u = p * sqrt(n) / (n_child + 1) * (log((n + c2 + 1) / c2) + c1) with the same normalised Q term."""
    kind = TZ_SEL_MUZERO_PUCT

    def __init__(self, c1: float = 1.25, c2: float = 19652, epsilon: float = 1e-8, q_transform=normalize_q_values):
        super().__init__(epsilon=epsilon)
        _check_transform(q_transform)
        self.c1 = c1
        self.c2 = c2
        self.q_transform = q_transform

    def get_config(self) -> Dict:
        return {"c1": self.c1, "c2": self.c2, "q_transform": self.q_transform.__name__, **super().get_config()}

    def kernel_params(self) -> Dict:
        return dict(selector=self.kind, c=1.0, c1=self.c1, c2=float(self.c2), epsilon=self.epsilon)
