"""AlphaZero root-noise mixin -- mirrors core/evaluators/alphazero.py of the reference.

The Dirichlet mixing and masked re-normalisation (alphazero.py:57-76) is host-framework math (PyTorch here,
XLA in the reference); only the final `update_root_node` + `set_root` write is a kernel (tz_set_root).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from .mcts import MCTS
from .trees import MCTSTree
from .types import StepMetadata


class _AlphaZero:
    """alphazero.py:12-81"""

    def __init__(self, dirichlet_alpha: float = 0.3, dirichlet_epsilon: float = 0.25, **kwargs):
        super().__init__(**kwargs)
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_epsilon = dirichlet_epsilon

    def get_config(self) -> Dict:
        """alphazero.py:34-40"""
        return {
            "dirichlet_alpha": self.dirichlet_alpha,
            "dirichlet_epsilon": self.dirichlet_epsilon,
            **super().get_config()  # pylint: disable=no-member
        }

    def update_root(self, key, tree: MCTSTree, root_embedding: Any, params: Any, root_metadata: StepMetadata,
                    dirichlet_noise: Optional[torch.Tensor] = None, **kwargs) -> MCTSTree:
        """alphazero.py:43-81.  `dirichlet_noise` (B,F) replaces the draw from `key` (used by parity tests)."""
        root_policy_logits, root_value = self.eval_fn(root_embedding, params, key)  # pylint: disable=no-member
        root_policy = torch.softmax(root_policy_logits, dim=-1)
        B, F = root_policy.shape
        if dirichlet_noise is None:
            gen = key if isinstance(key, torch.Generator) else None
            alpha = torch.full((B, F), self.dirichlet_alpha, dtype=root_policy.dtype, device=root_policy.device)
            dirichlet_noise = torch._sample_dirichlet(alpha, generator=gen)
        noisy_policy = ((1 - self.dirichlet_epsilon) * root_policy) + (self.dirichlet_epsilon * dirichlet_noise)
        finfo = torch.finfo(noisy_policy.dtype)
        new_logits = torch.log(torch.clamp(noisy_policy, min=finfo.tiny))
        policy = torch.where(root_metadata.action_mask.bool(), new_logits, finfo.min)
        renorm_policy = torch.softmax(policy, dim=-1)
        return self._set_root(tree, renorm_policy, root_value, root_embedding)  # pylint: disable=no-member


class AlphaZero(MCTS):
    """alphazero.py:84-98: `AlphaZero(WeightedMCTS)(...)` builds a class that extends the given MCTS backend."""

    def __new__(cls, base_type: type = MCTS):
        assert issubclass(base_type, MCTS)
        cls_type = type("AlphaZero", (_AlphaZero, base_type), {})
        cls_type.__name__ = f'AlphaZero({base_type.__name__})'
        return cls_type
