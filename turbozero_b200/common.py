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
    done = terminated | truncated
    if reset:
        eval_state = evaluator.step(output.eval_state, output.action, reset_mask=done)
        if env_init_fn is not None:
            env_state, env_state_metadata = env_init_fn(key, done, env_state, env_state_metadata)
    else:
        eval_state = evaluator.step(output.eval_state, output.action, keep_mask=done)  # `lambda s: s`, common.py:91
    output = output.replace(eval_state=eval_state)
    return output, env_state, env_state_metadata, terminated, truncated, rewards
