"""ORACLE / TEST INFRASTRUCTURE -- not a product path (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this).  NumPy restatement of the reference's episode replay buffer, core/memory/replay_memory.py, batched over
envs the way Trainer.collect applies it under vmap (core/training/train.py:271-347).  Pinned against the reference's own
source run through oracle/jaxshim (tests/golden/replay_*.npz, tests/test_replay_oracle.py).

State layout = ReplayBufferState (replay_memory.py:24-41) with the leading env axis init() gives it (:186-206):
next_idx [B] i32, episode_start_idx [B] i32, populated [B,cap] bool, has_reward [B,cap] bool, buffer leaves [B,cap,...].
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np


@dataclass
class ReplayState:
    next_idx: np.ndarray
    episode_start_idx: np.ndarray
    populated: np.ndarray
    has_reward: np.ndarray
    buffer: Dict[str, np.ndarray]  # BaseExperience fields (replay_memory.py:8-21), "reward" among them

    def copy(self) -> "ReplayState":
        return ReplayState(self.next_idx.copy(), self.episode_start_idx.copy(), self.populated.copy(), self.has_reward.copy(),
                           {k: v.copy() for k, v in self.buffer.items()})


def init(batch_size: int, capacity: int, template: Dict[str, np.ndarray]) -> ReplayState:
    """replay_memory.py:186-206: indices 0, buffer zeros, populated False, has_reward True."""
    return ReplayState(
        next_idx=np.zeros((batch_size,), np.int32),
        episode_start_idx=np.zeros((batch_size,), np.int32),
        populated=np.zeros((batch_size, capacity), bool),
        has_reward=np.ones((batch_size, capacity), bool),
        buffer={k: np.zeros((batch_size, capacity, *np.shape(v)), np.asarray(v).dtype) for k, v in template.items()},
    )


def add_experience(s: ReplayState, exp: Dict[str, np.ndarray], capacity: int) -> None:
    """replay_memory.py:65-84, every env: row next_idx <- experience; populated True, has_reward False there;
    next_idx = (next_idx + 1) % capacity."""
    b = np.arange(s.next_idx.shape[0])
    idx = s.next_idx.copy()
    for k, v in exp.items():
        s.buffer[k][b, idx] = v
    s.populated[b, idx] = True
    s.has_reward[b, idx] = False
    s.next_idx = ((idx + 1) % capacity).astype(np.int32)


def assign_rewards(s: ReplayState, reward: np.ndarray, mask: np.ndarray) -> None:
    """replay_memory.py:87-107 under `lax.cond(terminated, ...)` (train.py:327-332): where mask[b], every slot WITHOUT a
    reward gets reward[b] (populated or not), all slots are marked rewarded, episode_start_idx = next_idx."""
    m = mask.astype(bool)
    fill = m[:, None] & ~s.has_reward
    s.buffer["reward"] = np.where(fill[..., None], reward[:, None, :], s.buffer["reward"]).astype(s.buffer["reward"].dtype)
    s.has_reward = np.where(m[:, None], True, s.has_reward)
    s.episode_start_idx = np.where(m, s.next_idx, s.episode_start_idx).astype(np.int32)


def truncate(s: ReplayState, mask: np.ndarray) -> None:
    """replay_memory.py:110-135 under `lax.cond(truncated, ...)` (train.py:334-339): where mask[b], next_idx goes back to
    episode_start_idx, slots without a reward are un-populated, all slots are marked rewarded."""
    m = mask.astype(bool)
    s.populated = np.where(m[:, None] & ~s.has_reward, False, s.populated)
    s.has_reward = np.where(m[:, None], True, s.has_reward)
    s.next_idx = np.where(m, s.episode_start_idx, s.next_idx).astype(np.int32)


def collect_update(s: ReplayState, experiences, reward: np.ndarray, terminated: np.ndarray, truncated: np.ndarray,
                   capacity: int) -> None:
    """The buffer half of Trainer.collect (train.py:300-340): add_experience for the step (and one per transform),
    then assign_rewards where terminated, then truncate where truncated."""
    for exp in experiences:
        add_experience(s, exp, capacity)
    assign_rewards(s, reward, terminated)
    truncate(s, truncated)


def sample_indices(s: ReplayState, gumbel: np.ndarray, sample_size: int) -> np.ndarray:
    """replay_memory.py:157-169 over ONE device partition's (B, cap) block flattened: weights = populated & has_reward;
    jax.random.choice(replace=False, p = w / sum(w)) = argsort(-gumbel - log(p))[:sample_size] (stable)."""
    w = (s.populated & s.has_reward).reshape(-1).astype(np.float32)
    p = (w / np.float32(w.sum(dtype=np.float32))).astype(np.float32)
    logp = _logf(p)
    g = (-gumbel.astype(np.float32) - logp).astype(np.float32)
    return np.argsort(g, kind="stable")[:sample_size].astype(np.int64)


def _logf(p: np.ndarray) -> np.ndarray:
    """log of 0/1-weights normalised by their count: -inf for 0, log(1/n) otherwise (float32)."""
    from . import mcts_numpy as M  # tz_logf: this path's deterministic float32 log (include/tz_math.h)

    out = np.full(p.shape, -np.inf, np.float32)
    nz = p > 0
    out[nz] = M.tz_logf(p[nz])
    return out


def sample(s: ReplayState, gumbel: np.ndarray, sample_size: int) -> Dict[str, np.ndarray]:
    """replay_memory.py:171-181: rows at the sampled (batch, item) indices, for every buffer leaf."""
    idx = sample_indices(s, gumbel, sample_size)
    cap = s.populated.shape[1]
    bi, ii = idx // cap, idx % cap
    return {k: v[bi, ii] for k, v in s.buffer.items()}
