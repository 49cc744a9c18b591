"""Plug-in signatures of the search path -- mirrors core/types.py:11-32 of the reference.

Everything here is BATCHED: where the reference writes a per-environment function and lets `jax.vmap`
(core/training/train.py:613) add the env axis, this package passes tensors with a leading batch axis B.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any, Callable, Tuple

import torch


@dataclass(frozen=True)
class StepMetadata:
    """core/types.py:11-25.
    - `rewards`: (B, num_players) rewards received by the players
    - `action_mask`: (B, F) mask of valid actions
    - `terminated`: (B,) whether the environment is terminated
    - `cur_player_id`: (B,) current player id
    - `step`: (B,) step number
    """
    rewards: torch.Tensor
    action_mask: torch.Tensor
    terminated: torch.Tensor
    cur_player_id: torch.Tensor
    step: torch.Tensor

    def replace(self, **kw) -> "StepMetadata":
        return replace(self, **kw)


ArrayTree = Any
Params = Any
EnvStepFn = Callable[[ArrayTree, torch.Tensor], Tuple[ArrayTree, StepMetadata]]  # core/types.py:27
EnvInitFn = Callable[[Any], Tuple[ArrayTree, StepMetadata]]  # core/types.py:28
EvalFn = Callable[[ArrayTree, Params, Any], Tuple[torch.Tensor, torch.Tensor]]  # core/types.py:31
