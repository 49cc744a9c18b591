"""Episode glue around the evaluator -- mirrors core/common.py:12-103 of the reference (batched)."""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import torch
from torch.utils import _pytree as pytree

from .evaluator import EvalOutput, Evaluator
from .types import EnvInitFn, EnvStepFn, StepMetadata


def partition(data: Any, num_partitions: int) -> Any:
    """common.py:12-29: (N, ...) -> (num_partitions, N // num_partitions, ...) on every leaf."""
    return pytree.tree_map(lambda x: x.reshape(num_partitions, x.shape[0] // num_partitions, *x.shape[1:]), data)


def shard_slice(batch_size: int, rank: int, world_size: int) -> slice:
    """The contiguous env slice rank `rank` owns: row `rank` of `partition(..., world_size)` (common.py:26-29,
    core/training/train.py:204-217 requires divisibility)."""
    if batch_size % world_size != 0:
        raise ValueError(f"batch_size {batch_size} must be divisible by the number of devices {world_size}")
    per = batch_size // world_size
    return slice(rank * per, (rank + 1) * per)


def merge_topk(scores: torch.Tensor, global_index: torch.Tensor, k: int, per_rank: int, group=None):
    """Global `k` smallest of per-rank candidate lists (each rank passes its own k best, ascending, with indices offset by
    rank * per_rank): all-gather the candidates, sort by (score, global index) -- the stable order a single argsort over the
    concatenated blocks would give -- and return (owner rank, index inside the owner's block) of the winners.
    Used by EpisodeReplayBuffer.sample across ranks (core/memory/replay_memory.py:157-169 samples over the device axis)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    pad = k - scores.numel()
    if pad > 0:  # fewer than k slots on this rank
        scores = torch.cat([scores, torch.full((pad,), float("inf"), dtype=scores.dtype, device=scores.device)])
        global_index = torch.cat([global_index, torch.full((pad,), torch.iinfo(torch.int64).max, dtype=torch.int64,
                                                           device=global_index.device)])
    all_s = [torch.empty_like(scores) for _ in range(world)]
    all_i = [torch.empty_like(global_index) for _ in range(world)]
    dist.all_gather(all_s, scores.contiguous(), group=group)
    dist.all_gather(all_i, global_index.contiguous(), group=group)
    s, i = torch.cat(all_s), torch.cat(all_i)
    by_index = torch.sort(i, stable=True).indices            # candidates in global-index order ...
    pick = by_index[torch.sort(s[by_index], stable=True).indices[:k]]  # ... then a stable sort by score
    win = i[pick]
    return win // per_rank, win % per_rank


def step_env_and_evaluator(
    key,
    env_state: Any,
    env_state_metadata: StepMetadata,
    eval_state: Any,
    params: Any,
    evaluator: Evaluator,
    env_step_fn: EnvStepFn,
    env_init_fn: Optional[Callable[[Any, torch.Tensor, Any, StepMetadata], Tuple[Any, StepMetadata]]],
    max_steps: int,
    reset: bool = True,
    **evaluate_kwargs,
) -> Tuple[EvalOutput, Any, StepMetadata, torch.Tensor, torch.Tensor, torch.Tensor]:
    """common.py:32-103, batched.

    `env_init_fn(key, mask, env_state, metadata) -> (env_state, metadata)` re-initialises the envs flagged in
    `mask` (the batched form of `lax.cond(terminated | truncated, env_init_fn(key), identity)`, common.py:95-100).
    The evaluator's reset-vs-step select (common.py:89-94) is one fused re-root launch.
    """
    output = evaluator.evaluate(key=key, eval_state=eval_state, env_state=env_state, root_metadata=env_state_metadata,
                                params=params, env_step_fn=env_step_fn, **evaluate_kwargs)
    env_state, env_state_metadata = env_step_fn(env_state, output.action)
    terminated = env_state_metadata.terminated.bool()
    truncated = env_state_metadata.step > max_steps  # strict, common.py:86
    rewards = env_state_metadata.rewards
    if reset and env_init_fn is not None:
        rewards = rewards.clone()  # an env_init_fn that resets metadata in place must not change what is returned
    done = terminated | truncated
    if reset:
        eval_state = evaluator.step(output.eval_state, output.action, reset_mask=done)
        if env_init_fn is not None:
            env_state, env_state_metadata = env_init_fn(key, done, env_state, env_state_metadata)
    else:
        eval_state = evaluator.step(output.eval_state, output.action, keep_mask=done)  # `lambda s: s`, common.py:91
    output = output.replace(eval_state=eval_state)
    return output, env_state, env_state_metadata, terminated, truncated, rewards


# ----------------------------------------------------------------------------------------------------------------------
# two evaluators playing each other: core/common.py:104-367
# ----------------------------------------------------------------------------------------------------------------------
from dataclasses import dataclass as _dataclass, replace as _replace  # noqa: E402
from typing import List  # noqa: E402


@_dataclass(frozen=True)
class TwoPlayerGameState:
    """common.py:104-126, batched: every field carries a leading game axis G."""
    key: Any
    env_state: Any
    env_state_metadata: StepMetadata
    p1_eval_state: Any
    p2_eval_state: Any
    p1_value_estimate: torch.Tensor  # (G,)
    p2_value_estimate: torch.Tensor  # (G,)
    outcomes: torch.Tensor           # (G, 2)
    completed: torch.Tensor          # (G,) bool

    def replace(self, **kw) -> "TwoPlayerGameState":
        return _replace(self, **kw)


@_dataclass(frozen=True)
class GameFrame:
    """common.py:129-143 (rendering record).  `env_state` is a copy taken when the frame was made."""
    env_state: Any
    p1_value_estimate: torch.Tensor
    p2_value_estimate: torch.Tensor
    completed: torch.Tensor
    outcomes: torch.Tensor


def _select_tree(mask: torch.Tensor, a: Any, b: Any) -> Any:
    """where(mask, a, b) on every leaf (mask over the leading game axis)."""
    def pick(x, y):
        if not isinstance(x, torch.Tensor):
            return x
        m = mask.reshape(-1, *([1] * (x.dim() - 1)))
        return torch.where(m, x, y)
    return pytree.tree_map(pick, a, b)


def _clone_tree(tree: Any) -> Any:
    return pytree.tree_map(lambda t: t.clone() if isinstance(t, torch.Tensor) else t, tree)


def two_player_game_step(state: TwoPlayerGameState, p1_evaluator: Evaluator, p2_evaluator: Evaluator, params: Any,
                         env_step_fn: EnvStepFn, env_init_fn, use_p1: bool, max_steps: int, *, return_action: bool = False,
                         **evaluate_kwargs) -> TwoPlayerGameState:
    """common.py:146-232 for a batch of games in which the SAME evaluator is to move (`use_p1`).

    Games already `completed` are not stepped in the reference (`lax.cond(state.completed, identity, ...)`,
    common.py:305-317); here the whole batch is stepped and the completed games' env state, value estimates, outcomes
    are kept by a select -- their trees are no longer read by anything."""
    if use_p1:
        active_evaluator, other_evaluator = p1_evaluator, p2_evaluator
        active_eval_state, other_eval_state = state.p1_eval_state, state.p2_eval_state
    else:
        active_evaluator, other_evaluator = p2_evaluator, p1_evaluator
        active_eval_state, other_eval_state = state.p2_eval_state, state.p1_eval_state
    done_before = state.completed
    env_before = _clone_tree(state.env_state)  # the user's env may step in place
    output, env_state, env_state_metadata, terminated, truncated, rewards = step_env_and_evaluator(
        key=state.key, env_state=state.env_state, env_state_metadata=state.env_state_metadata, eval_state=active_eval_state,
        params=params, evaluator=active_evaluator, env_step_fn=env_step_fn, env_init_fn=env_init_fn, max_steps=max_steps,
        reset=False, **evaluate_kwargs)  # common.py:181-192
    done = terminated | truncated
    active_eval_state = output.eval_state
    active_value = active_evaluator.get_value(active_eval_state)
    active_value = torch.where(done, active_value, active_evaluator.discount * active_value)  # common.py:197-202
    other_eval_state = other_evaluator.step(other_eval_state, output.action)  # common.py:204 (absent child => empty tree)
    other_value = other_evaluator.get_value(other_eval_state)
    if use_p1:
        p1_state, p2_state, p1_v, p2_v = active_eval_state, other_eval_state, active_value, other_value
    else:
        p1_state, p2_state, p1_v, p2_v = other_eval_state, active_eval_state, other_value, active_value
    md = env_state_metadata
    md_old = state.env_state_metadata
    keep = done_before
    new_state = state.replace(
        env_state=_select_tree(keep, env_before, env_state),
        env_state_metadata=StepMetadata(**{f: torch.where(keep.reshape(-1, *([1] * (getattr(md, f).dim() - 1))), getattr(md_old, f).to(getattr(md, f).dtype), getattr(md, f))
                                           for f in ("rewards", "action_mask", "terminated", "cur_player_id", "step")}),
        p1_eval_state=p1_state, p2_eval_state=p2_state,
        p1_value_estimate=torch.where(keep, state.p1_value_estimate, p1_v.clone()),
        p2_value_estimate=torch.where(keep, state.p2_value_estimate, p2_v.clone()),
        outcomes=torch.where((done & ~done_before).unsqueeze(-1), rewards.to(state.outcomes.dtype), state.outcomes),  # :222-226
        completed=done_before | done,
    )
    return (new_state, output.action) if return_action else new_state


def two_player_game(key, evaluator_1: Evaluator, evaluator_2: Evaluator, params_1: Any, params_2: Any, env_step_fn: EnvStepFn,
                    env_init_fn, max_steps: int, *, num_games: int, p1_first: bool, template_embedding: Any = None,
                    frames: bool = False, **evaluate_kwargs):
    """common.py:235-367 for `num_games` games played side by side in which evaluator 1 moves first iff `p1_first`
    (the reference draws that per game, common.py:276-283; a caller that wants the mix runs the two groups -- games are
    independent).  `env_init_fn(key, num_games) -> (env_state, metadata)`.

    Returns (outcomes (G, 2) ordered [evaluator 1, evaluator 2], final TwoPlayerGameState, frames or None, p_ids (G, 2))."""
    env_state, metadata = env_init_fn(key, num_games)
    emb = env_state if template_embedding is None else template_embedding
    one = pytree.tree_map(lambda t: t[0], emb)
    dev = metadata.cur_player_id.device
    p1_eval_state = evaluator_1.init_batched(num_games, one, device=dev)
    p2_eval_state = evaluator_2.init_batched(num_games, one, device=dev)
    cur = metadata.cur_player_id.long()
    p1_id, p2_id = (cur, 1 - cur) if p1_first else (1 - cur, cur)  # common.py:278-283
    state = TwoPlayerGameState(
        key=key, env_state=env_state, env_state_metadata=metadata, p1_eval_state=p1_eval_state, p2_eval_state=p2_eval_state,
        p1_value_estimate=torch.zeros((num_games,), dtype=torch.float32, device=dev),
        p2_value_estimate=torch.zeros((num_games,), dtype=torch.float32, device=dev),
        outcomes=torch.zeros((num_games, 2), dtype=torch.float32, device=dev),
        completed=torch.zeros((num_games,), dtype=torch.bool, device=dev))

    def frame(s):
        return GameFrame(env_state=_clone_tree(s.env_state), p1_value_estimate=s.p1_value_estimate.clone(),
                         p2_value_estimate=s.p2_value_estimate.clone(), completed=s.completed.clone(), outcomes=s.outcomes.clone())

    out_frames: Optional[List[GameFrame]] = [frame(state)] if frames else None
    for _ in range(max_steps // 2):  # common.py:303-355: a turn for each player per scan step
        for first_half in (True, False):
            use_p1 = p1_first == first_half
            state = two_player_game_step(state, evaluator_1, evaluator_2, params_1 if use_p1 else params_2, env_step_fn,
                                            env_init_fn, use_p1, max_steps, **evaluate_kwargs)
            if frames:
                out_frames.append(frame(state))
    outcomes = torch.stack([state.outcomes.gather(1, p1_id.unsqueeze(1)).squeeze(1),
                            state.outcomes.gather(1, p2_id.unsqueeze(1)).squeeze(1)], dim=1)  # common.py:367
    return outcomes, state, out_frames, torch.stack([p1_id, p2_id], dim=1)
