#!/usr/bin/env python
"""Generates tests/golden/replay_*.npz by executing the reference's UNMODIFIED core/memory/replay_memory.py on the NumPy
emulation of the jax API (oracle/jaxshim): per env, the buffer half of Trainer.collect (core/training/train.py:300-340:
add_experience [+ one per transform] -> cond(terminated, assign_rewards) -> cond(truncated, truncate)), then one
EpisodeReplayBuffer.sample over the (1, B, capacity) state.  Inputs (experiences, masks, gumbel noise) are stored in the
fixture, so no RNG stream has to be reproducible.

    python tests/golden/make_golden_replay.py            # regenerate (needs /root/reference)
    python tests/golden/make_golden_replay.py --check    # regenerate in memory and compare with the committed files
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_via_shim as R  # noqa: E402

# name -> (B, capacity, players, F, obs shape, steps, experiences per step, P(terminated), P(truncated), sample size, seed)
CASES = {
    "replay_small": (5, 7, 2, 4, (2, 3), 30, 1, 0.15, 0.05, 6, 1),
    "replay_transforms_wrap": (4, 10, 2, 9, (3, 3, 2), 40, 2, 0.25, 0.05, 5, 2),  # 2 rows per step, ring wraps inside episodes
    "replay_single_player": (6, 16, 1, 4, (4, 4), 50, 1, 0.08, 0.02, 12, 3),      # 2048-like: one reward column
}


def make_inputs(case):
    B, cap, P, F, obs, steps, n_exp, p_term, p_trunc, S, seed = case
    rng = np.random.default_rng(seed)
    x = {
        "observation_nn": rng.standard_normal((steps, n_exp, B, *obs)).astype(np.float32),
        "policy_weights": rng.dirichlet([0.5] * F, size=(steps, n_exp, B)).astype(np.float32),
        "policy_mask": rng.random((steps, n_exp, B, F)) < 0.8,
        "cur_player_id": rng.integers(0, max(P, 1), size=(steps, n_exp, B)).astype(np.int32),
        "rewards": rng.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(steps, B, P)).astype(np.float32),
        "terminated": rng.random((steps, B)) < p_term,
        "truncated": rng.random((steps, B)) < p_trunc,
        "gumbel": rng.gumbel(size=(B * cap,)).astype(np.float32),
    }
    return x


def run_reference(case, x):
    B, cap, P, F, obs, steps, n_exp, p_term, p_trunc, S, seed = case
    R.load_reference()
    sys.path.insert(0, R.SHIM_ROOT)
    sys.path.insert(0, R.REFERENCE_ROOT)
    try:
        rm = importlib.import_module("core.memory.replay_memory")
        import jax
        import jax.numpy as jnp
    finally:
        sys.path.remove(R.SHIM_ROOT)
        sys.path.remove(R.REFERENCE_ROOT)
    buf = rm.EpisodeReplayBuffer(capacity=cap)
    tmpl = rm.BaseExperience(reward=jnp.zeros((P,), dtype=jnp.float32), policy_weights=jnp.zeros((F,), dtype=jnp.float32),
                             policy_mask=jnp.zeros((F,), dtype=jnp.bool_), observation_nn=jnp.zeros(obs, dtype=jnp.float32),
                             cur_player_id=jnp.zeros((), dtype=jnp.int32))
    batched = buf.init(B, tmpl)
    envs = [jax.tree_util.tree_map(lambda a, b=b: a[b], batched) for b in range(B)]  # what vmap hands each env
    for t in range(steps):
        for b in range(B):
            st = envs[b]
            for e in range(n_exp):  # train.py:300-325
                st = buf.add_experience(st, rm.BaseExperience(
                    observation_nn=jnp.array(x["observation_nn"][t, e, b]), policy_mask=jnp.array(x["policy_mask"][t, e, b]),
                    policy_weights=jnp.array(x["policy_weights"][t, e, b]),
                    reward=jnp.empty_like(jnp.array(x["rewards"][t, b])), cur_player_id=jnp.array(x["cur_player_id"][t, e, b])))
            rew = jnp.array(x["rewards"][t, b])
            st = jax.lax.cond(bool(x["terminated"][t, b]), lambda s: buf.assign_rewards(s, rew), lambda s: s, st)  # :327-332
            st = jax.lax.cond(bool(x["truncated"][t, b]), buf.truncate, lambda s: s, st)                          # :334-339
            envs[b] = st
    out = {
        "next_idx": np.array([int(e.next_idx) for e in envs], np.int32),
        "episode_start_idx": np.array([int(e.episode_start_idx) for e in envs], np.int32),
        "populated": np.stack([np.asarray(e.populated).astype(bool) for e in envs]),
        "has_reward": np.stack([np.asarray(e.has_reward).astype(bool) for e in envs]),
    }
    for f in ("reward", "policy_weights", "policy_mask", "observation_nn", "cur_player_id"):
        out["buf_" + f] = np.stack([np.asarray(getattr(e.buffer, f)) for e in envs])
    # sample over (devices=1, B, cap): replay_memory.py:137-183 with the gumbel noise served from the fixture
    class Tape(jax.random.Tape):
        def gumbel(self, key, shape):
            assert shape == (B * cap,)
            return x["gumbel"]
    jax.random.install_tape(Tape())
    state = rm.ReplayBufferState(
        next_idx=jnp.array(out["next_idx"][None]), episode_start_idx=jnp.array(out["episode_start_idx"][None]),
        buffer=rm.BaseExperience(**{f: jnp.array(out["buf_" + f][None]) for f in
                                    ("reward", "policy_weights", "policy_mask", "observation_nn", "cur_player_id")}),
        populated=jnp.array(out["populated"][None]), has_reward=jnp.array(out["has_reward"][None]))
    smp = buf.sample(state, jax.random.PRNGKey(0), S)
    for f in ("reward", "policy_weights", "policy_mask", "observation_nn", "cur_player_id"):
        out["sample_" + f] = np.asarray(getattr(smp, f))
    return out


def main():
    check = "--check" in sys.argv
    for name, case in CASES.items():
        x = make_inputs(case)
        out = run_reference(case, x)
        blob = {}
        blob.update({"in_" + k: v for k, v in x.items()})
        blob.update({"ref_" + k: v for k, v in out.items()})
        path = os.path.join(HERE, name + ".npz")
        if check:
            old = np.load(path)
            for k, v in blob.items():
                assert np.array_equal(old[k], v), f"{name}: {k} differs from the committed fixture"
            print(f"{name}: matches")
        else:
            np.savez_compressed(path, **blob)
            print(f"{name}: wrote {os.path.getsize(path)} bytes; valid rows {int((out['populated'] & out['has_reward']).sum())}")


if __name__ == "__main__":
    main()
