"""One self-play collection step -- mirrors Trainer.collect of the reference (core/training/train.py:271-347), batched:
step the envs with the evaluator's action (step_env_and_evaluator), store the step's experience (and one per data
transform) in the episode replay buffer, assign rewards where the episode ended, drop the episode where it was cut off.
Search and buffer updates are CUDA kernels (include/tz_abi.h, include/tz_replay.h); the env, the network and the
observation / transform functions are the user's, exactly as in the reference."""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Any, Callable, Optional, Sequence

import torch

from .common import step_env_and_evaluator
from .evaluator import Evaluator
from .replay_memory import BaseExperience, EpisodeReplayBuffer, ReplayBufferState
from .types import StepMetadata


@dataclass(frozen=True)
class CollectionState:
    """train.py:26-37"""
    eval_state: Any
    env_state: Any
    buffer_state: ReplayBufferState
    metadata: StepMetadata

    def replace(self, **kw) -> "CollectionState":
        return replace(self, **kw)


def collect(key, state: CollectionState, params, *, evaluator: Evaluator, env_step_fn, env_init_fn, max_steps: int,
            memory_buffer: EpisodeReplayBuffer, state_to_nn_input_fn: Callable[[Any], torch.Tensor],
            transform_fns: Sequence[Callable] = (), **evaluate_kwargs) -> CollectionState:
    """train.py:271-347.  The experience describes the position the search ran on, so observation, action mask and player
    are taken from `state` BEFORE the env is stepped (user envs that step in place are fine: they are read first)."""
    md = state.metadata
    obs = state_to_nn_input_fn(state.env_state)
    mask, player = md.action_mask, md.cur_player_id
    pre = [(obs.clone(), mask.clone(), player.clone())]
    # the transforms need the search's policy weights but the PRE-step env state: keep a copy only if there are any
    env_before = _clone_tree(state.env_state) if transform_fns else None
    out, env_state, metadata, terminated, truncated, rewards = step_env_and_evaluator(
        key=key, env_state=state.env_state, env_state_metadata=md, eval_state=state.eval_state, params=params,
        evaluator=evaluator, env_step_fn=env_step_fn, env_init_fn=env_init_fn, max_steps=max_steps, **evaluate_kwargs)
    zeros = torch.zeros_like(rewards)  # train.py:307 `jnp.empty_like(rewards)`: XLA materialises it as zeros
    exps = [BaseExperience(observation_nn=pre[0][0], policy_mask=pre[0][1], policy_weights=out.policy_weights, reward=zeros,
                           cur_player_id=pre[0][2])]
    for fn in transform_fns:  # train.py:311-325
        t_mask, t_pw, t_env = fn(pre[0][1], out.policy_weights, env_before)
        exps.append(BaseExperience(observation_nn=state_to_nn_input_fn(t_env), policy_mask=t_mask, policy_weights=t_pw,
                                   reward=zeros, cur_player_id=pre[0][2]))
    buffer_state = memory_buffer.collect_update(state.buffer_state, exps, rewards, terminated, truncated)  # :300-340
    return state.replace(eval_state=out.eval_state, env_state=env_state, buffer_state=buffer_state, metadata=metadata)


def _clone_tree(tree):
    from torch.utils import _pytree as pytree

    return pytree.tree_map(lambda t: t.clone() if isinstance(t, torch.Tensor) else t, tree)
