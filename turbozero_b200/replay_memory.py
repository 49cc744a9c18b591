"""Episode replay buffer on sm_100a kernels -- the drop-in for core/memory/replay_memory.py of the reference, batched
over envs the way Trainer.collect applies it under vmap (core/training/train.py:271-347).

Same names and meaning (`BaseExperience`, `ReplayBufferState`, `EpisodeReplayBuffer.init / add_experience /
assign_rewards / truncate / sample / get_config`); what differs, by design:
 * every method takes the whole env batch; the `lax.cond(terminated, ...)` / `lax.cond(truncated, ...)` of the caller
   are per-env masks, and `collect_update` does the three buffer updates of one collection step in ONE launch;
 * buffers are updated in place and the same state object is returned;
 * `sample` takes the Gumbel noise jax.random.choice would draw as an explicit input when bit-parity is wanted.
There is no CPU / PyTorch fallback for the buffer updates: without libtz_b200.so this module raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields
from typing import Any, Dict, List, Optional, Sequence

import torch
from torch.utils import _pytree as pytree

from . import _abi
from .trees import _on_device, _stream_ptr


@dataclass(frozen=True)
class BaseExperience:
    """replay_memory.py:8-21"""
    reward: torch.Tensor
    policy_weights: torch.Tensor
    policy_mask: torch.Tensor
    observation_nn: torch.Tensor
    cur_player_id: torch.Tensor


@dataclass
class ReplayBufferState:
    """replay_memory.py:24-41 with the leading env axis of EpisodeReplayBuffer.init (:186-206)."""
    next_idx: torch.Tensor           # (B,) int32
    episode_start_idx: torch.Tensor  # (B,) int32
    buffer: Any                      # experience pytree, leaves (B, capacity, ...)
    populated: torch.Tensor          # (B, capacity) bool
    has_reward: torch.Tensor         # (B, capacity) bool
    _leaves: List[torch.Tensor] = None
    _struct: Any = None

    def struct(self) -> _abi.TzReplay:
        return self._struct


def _flatten(exp) -> List[torch.Tensor]:
    if hasattr(exp, "__dataclass_fields__"):
        return [getattr(exp, f.name) for f in sorted(fields(exp), key=lambda f: f.name)]  # chex order: sorted by name
    return pytree.tree_leaves(exp)


def _row_view(t: torch.Tensor, lead: int) -> torch.Tensor:
    """(lead dims..., row...) tensor as contiguous bytes-per-row bookkeeping: returns the tensor made contiguous."""
    return t if t.is_contiguous() else t.contiguous()


class EpisodeReplayBuffer:
    """replay_memory.py:44-206."""

    def __init__(self, capacity: int):
        self.capacity = capacity

    def get_config(self) -> Dict:
        """replay_memory.py:60-63"""
        return {"capacity": self.capacity}

    # ------------------------------------------------------------------------------------------------
    def init(self, batch_size: int, template_experience, device="cuda") -> ReplayBufferState:
        """replay_memory.py:186-206: zeros everywhere, populated False, has_reward True."""
        _abi.lib()  # raises if the CUDA library is missing
        dev = torch.device(device)
        tl = _flatten(template_experience)
        names = ([f.name for f in sorted(fields(template_experience), key=lambda f: f.name)]
                 if hasattr(template_experience, "__dataclass_fields__") else None)
        leaves = [torch.zeros((batch_size, self.capacity, *t.shape), dtype=t.dtype, device=dev) for t in tl]
        if names is not None:
            buffer = type(template_experience)(**dict(zip(names, leaves)))
            reward_leaf = names.index("reward")
        else:
            buffer = pytree.tree_unflatten(leaves, pytree.tree_structure(template_experience))
            reward_leaf = 0
        st = ReplayBufferState(
            next_idx=torch.zeros((batch_size,), dtype=torch.int32, device=dev),
            episode_start_idx=torch.zeros((batch_size,), dtype=torch.int32, device=dev),
            buffer=buffer,
            populated=torch.zeros((batch_size, self.capacity), dtype=torch.bool, device=dev),
            has_reward=torch.ones((batch_size, self.capacity), dtype=torch.bool, device=dev),
            _leaves=leaves)
        if leaves[reward_leaf].dtype != torch.float32:
            raise TypeError("the `reward` leaf must be float32")
        r = _abi.TzReplay()
        r.B, r.capacity, r.n_leaves, r.reward_leaf = batch_size, self.capacity, len(leaves), reward_leaf
        r.reward_dim = int(leaves[reward_leaf][0, 0].numel())
        r.next_idx, r.episode_start_idx = st.next_idx.data_ptr(), st.episode_start_idx.data_ptr()
        r.populated, r.has_reward = st.populated.data_ptr(), st.has_reward.data_ptr()
        for k, t in enumerate(leaves):
            r.leaf[k] = t.data_ptr()
            r.leaf_row_bytes[k] = int(t[0, 0].numel()) * t.element_size()
        st._struct = r
        return st

    # ------------------------------------------------------------------------------------------------
    def _exp_ptrs(self, state: ReplayBufferState, experiences: Sequence) -> "tuple[Any, list]":
        keep, ptrs = [], []
        for exp in experiences:
            for t, slot in zip(_flatten(exp), state._leaves):
                if tuple(t.shape) != (slot.shape[0], *slot.shape[2:]):
                    raise ValueError(f"experience leaf has shape {tuple(t.shape)}, expected {(slot.shape[0], *slot.shape[2:])}")
                t = t.to(dtype=slot.dtype).contiguous()
                keep.append(t)
                ptrs.append(t.data_ptr())
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        return arr, keep

    def collect_update(self, state: ReplayBufferState, experiences: Sequence, reward: Optional[torch.Tensor],
                       terminated: Optional[torch.Tensor], truncated: Optional[torch.Tensor]) -> ReplayBufferState:
        """The buffer half of Trainer.collect (train.py:300-340) in one launch: add every experience of the step, then
        assign_rewards where `terminated`, then truncate where `truncated`."""
        arr, keep = self._exp_ptrs(state, experiences)
        rew = None if reward is None else reward.to(torch.float32).contiguous()
        term = None if terminated is None else terminated.to(torch.uint8).contiguous()
        trunc = None if truncated is None else truncated.to(torch.uint8).contiguous()
        with _on_device(state.populated.device):
            _abi.check(_abi.lib().tz_replay_collect(
                C.byref(state._struct), len(experiences), arr, None if rew is None else rew.data_ptr(),
                None if term is None else term.data_ptr(), None if trunc is None else trunc.data_ptr(), _stream_ptr()),
                "tz_replay_collect")
        return state

    def add_experience(self, state: ReplayBufferState, experience) -> ReplayBufferState:
        """replay_memory.py:65-84 for every env."""
        return self.collect_update(state, [experience], None, None, None)

    def assign_rewards(self, state: ReplayBufferState, reward: torch.Tensor, mask: Optional[torch.Tensor] = None) -> ReplayBufferState:
        """replay_memory.py:87-107 for the envs in `mask` (all if None): the caller's lax.cond(terminated, ...)."""
        if mask is None:
            mask = torch.ones((state.next_idx.shape[0],), dtype=torch.uint8, device=state.next_idx.device)
        return self.collect_update(state, [], reward, mask, None)

    def truncate(self, state: ReplayBufferState, mask: Optional[torch.Tensor] = None) -> ReplayBufferState:
        """replay_memory.py:110-135 for the envs in `mask` (all if None): the caller's lax.cond(truncated, ...)."""
        if mask is None:
            mask = torch.ones((state.next_idx.shape[0],), dtype=torch.uint8, device=state.next_idx.device)
        return self.collect_update(state, [], None, None, mask)

    # ------------------------------------------------------------------------------------------------
    def sample_scores(self, state: ReplayBufferState, gumbel: torch.Tensor, group=None) -> torch.Tensor:
        """Keys of jax.random.choice(replace=False, p=mask/sum) (replay_memory.py:157-169): -gumbel - log(p), +inf where a
        slot cannot be sampled; (B*capacity,) float32.  p = 1 / (number of sampleable slots over all ranks)."""
        import torch.distributed as dist

        n = state.populated.numel()
        scores = torch.empty((n,), dtype=torch.float32, device=gumbel.device)
        n_valid = torch.empty((1,), dtype=torch.int32, device=gumbel.device)
        g = gumbel.to(torch.float32).contiguous()
        lib = _abi.lib()
        with _on_device(state.populated.device):
            _abi.check(lib.tz_replay_count_valid(C.byref(state._struct), n_valid.data_ptr(), _stream_ptr()), "tz_replay_count_valid")
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(n_valid, op=dist.ReduceOp.SUM, group=group)
            _abi.check(lib.tz_replay_sample_scores(C.byref(state._struct), g.data_ptr(), n_valid.data_ptr(), scores.data_ptr(),
                                                   _stream_ptr()), "tz_replay_sample_scores")
        return scores

    def gather(self, state: ReplayBufferState, flat_index: torch.Tensor):
        """replay_memory.py:171-181: rows at flat (env, item) indices of this device's block, for every leaf."""
        idx = flat_index.to(torch.int64).contiguous()
        n = int(idx.numel())
        outs = [torch.empty((n, *t.shape[2:]), dtype=t.dtype, device=t.device) for t in state._leaves]
        ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        with _on_device(state.populated.device):
            _abi.check(_abi.lib().tz_replay_gather(C.byref(state._struct), idx.data_ptr(), n, ptrs, _stream_ptr()), "tz_replay_gather")
        if hasattr(state.buffer, "__dataclass_fields__"):
            names = [f.name for f in sorted(fields(state.buffer), key=lambda f: f.name)]
            return type(state.buffer)(**dict(zip(names, outs)))
        return pytree.tree_unflatten(outs, pytree.tree_structure(state.buffer))

    def sample(self, state: ReplayBufferState, key, sample_size: int, gumbel: Optional[torch.Tensor] = None, group=None):
        """replay_memory.py:137-183.  Samples `sample_size` rows without replacement, uniformly over the populated and
        rewarded slots of ALL envs -- of all ranks when torch.distributed is initialised (the reference samples across
        its device axis too): every rank contributes its `sample_size` best keys, the global best are chosen from the
        gathered candidates, the owners gather the rows and the result is summed over ranks (disjoint ownership).
        `gumbel` (B*capacity,) pins the noise; otherwise it is drawn from `key` (a torch.Generator / int seed / None)."""
        dev = state.populated.device
        n = state.populated.numel()
        if gumbel is None:
            import torch.distributed as dist_

            rank_ = dist_.get_rank(group) if (dist_.is_available() and dist_.is_initialized()) else 0
            # every rank scores ITS OWN slots: the noise streams of the ranks must differ, or slot i gets the same key on
            # every rank and the draw is no longer uniform over the global buffer.  An int seed is folded with the rank; a
            # Generator is the caller's responsibility (seed it per rank).
            from .mcts import _generator

            if isinstance(key, torch.Generator):
                gen = key
            else:  # int seed, or None = a private stream derived from the process's global seed; both advance per draw
                gen = _generator(int(key if isinstance(key, int) else torch.initial_seed()) + 1000003 * rank_, dev)
            u = torch.rand((n,), dtype=torch.float32, device=dev, generator=gen).clamp_(min=torch.finfo(torch.float32).tiny)
            gumbel = -torch.log(-torch.log(u))
        scores = self.sample_scores(state, gumbel, group)
        order = torch.sort(scores, stable=True).indices[:sample_size]  # argsort(g)[:n_draws], ties by index
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return self.gather(state, order)
        from .common import merge_topk

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        owner, local = merge_topk(scores[order], order + rank * n, sample_size, n, group)
        mine = owner == rank
        rows = self.gather(state, torch.where(mine, local, torch.zeros_like(local)))
        leaves = _flatten(rows)
        # zero the rows other ranks own, then sum over ranks (bytes add up because ownership is disjoint).  All leaves travel as ONE
        # byte buffer -> one all-reduce per sample instead of one per leaf; masked_fill instead of boolean indexing: no host sync
        k = int(local.numel())
        views = [(t.view(torch.uint8) if t.dtype == torch.bool else t).reshape(k, -1).view(torch.uint8) for t in leaves]
        widths = [int(v.shape[1]) for v in views]
        packed = torch.cat(views, dim=1)
        packed.masked_fill_(~mine.unsqueeze(1), 0)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t, v, w in zip(leaves, views, widths):
            v.copy_(packed[:, off:off + w])
            off += w
        return rows
