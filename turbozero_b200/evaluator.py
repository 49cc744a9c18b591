"""Evaluator base class -- mirrors core/evaluators/evaluator.py:10-97 of the reference."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any, Dict

import torch


@dataclass(frozen=True)
class EvalOutput:
    """core/evaluators/evaluator.py:10-19.
    - `eval_state`: the updated internal state of the Evaluator
    - `action`: (B,) the action to take
    - `policy_weights`: (B, F) the policy weights assigned to each action
    """
    eval_state: Any
    action: torch.Tensor
    policy_weights: torch.Tensor

    def replace(self, **kw):
        return replace(self, **kw)


class Evaluator:
    """core/evaluators/evaluator.py:22-97."""

    def __init__(self, discount: float, *args, **kwargs):  # pylint: disable=unused-argument
        self.discount = discount

    def init(self, *args, **kwargs):
        raise NotImplementedError()

    def init_batched(self, batch_size: int, *args, **kwargs):
        """evaluator.py:42-45.  Subclasses allocate the batch natively instead of broadcasting one tree."""
        raise NotImplementedError()

    def reset(self, state):
        raise NotImplementedError()

    def evaluate(self, key, eval_state, env_state, **kwargs) -> EvalOutput:
        raise NotImplementedError()

    def step(self, state, action, reset_mask=None, keep_mask=None):  # pylint: disable=unused-argument
        """evaluator.py:62-72.  `reset_mask` / `keep_mask` (B,) carry the caller's per-env reset-vs-step select
        (core/common.py:89-94), which the batched `step_env_and_evaluator` folds into this call: envs in `reset_mask`
        get `reset`, envs in `keep_mask` keep their state.  A stateless evaluator ignores both."""
        return state

    def get_value(self, state) -> torch.Tensor:
        raise NotImplementedError()

    def get_config(self) -> Dict:
        """evaluator.py:95-97"""
        return {'discount': self.discount}
