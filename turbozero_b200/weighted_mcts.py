"""WeightedMCTS -- mirrors core/evaluators/mcts/weighted_mcts.py of the reference.

The node carries the extra `r` leaf (raw leaf value, weighted_mcts.py:14-17) and backpropagation is the
softmax-weighted child-Q backup of weighted_mcts.py:90-152, executed by the WEIGHTED instantiation of the
simulation kernel (csrc/tz_kernels.cu do_weighted_backprop).
"""
from __future__ import annotations

from typing import Dict

import torch

from .mcts import MCTS, _rand
from .trees import WeightedMCTSNode  # noqa: F401  (re-export, weighted_mcts.py:14)


class WeightedMCTS(MCTS):
    """weighted_mcts.py:20-152"""

    weighted = True

    def __init__(self, q_temperature: float = 1.0, *args, **kwargs):
        """- `q_temperature`: temperature applied to child q-values when backpropagating (weighted_mcts.py:25-33)"""
        super().__init__(*args, **kwargs)
        self.q_temperature = q_temperature

    def get_config(self) -> Dict:
        """weighted_mcts.py:35-40"""
        return {"q_temperature": self.q_temperature, **super().get_config()}

    def _backprop_noise(self, key, tree, backprop_noise, s):
        """weighted_mcts.py:123: only the q_temperature == 0 branch draws noise (one (F,) vector per simulation)."""
        if self.q_temperature > 0:
            return None
        if backprop_noise is not None:
            return backprop_noise[s].to(torch.float32).contiguous()
        return (_rand(key, (tree.batch_size, tree.branching_factor), tree.device) * self.tiebreak_noise).contiguous()
