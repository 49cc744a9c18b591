"""turbozero_b200 -- turbozero's batched MCTS hot path as hand-written sm_100a CUDA kernels behind the
reference's own Evaluator / MCTS API.  See DESIGN.md and INTEGRATION.md."""
from .action_selection import (MCTSActionSelector, MuZeroPUCTSelector, PUCTSelector, identity_q_values, normalize_q_values,
                               register_q_transform)
from .alphazero import AlphaZero
from .collect import CollectionState, collect
from .common import (GameFrame, TwoPlayerGameState, merge_topk, partition, shard_slice, step_env_and_evaluator,
                     two_player_game, two_player_game_step)
from .evaluator import EvalOutput, Evaluator
from .mcts import MCTS, MCTSOutput, TraversalState
from .replay_memory import BaseExperience, EpisodeReplayBuffer, ReplayBufferState
from .trees import MCTSNode, MCTSTree, Tree, WeightedMCTSNode, init_tree
from .types import StepMetadata
from .weighted_mcts import WeightedMCTS

__all__ = [
    "MCTS", "WeightedMCTS", "AlphaZero", "MCTSOutput", "TraversalState", "Evaluator", "EvalOutput",
    "MCTSActionSelector", "PUCTSelector", "MuZeroPUCTSelector", "normalize_q_values", "identity_q_values",
    "register_q_transform",
    "Tree", "MCTSTree", "MCTSNode", "WeightedMCTSNode", "init_tree", "StepMetadata",
    "partition", "shard_slice", "step_env_and_evaluator", "merge_topk",
    "TwoPlayerGameState", "GameFrame", "two_player_game_step", "two_player_game",
    "BaseExperience", "ReplayBufferState", "EpisodeReplayBuffer", "CollectionState", "collect",
]
