"""ORACLE (test infrastructure) -- NumPy restatement of the synthetic hash game (standin/include/tz_synth.h) and a
per-tree self-play driver that wires it to oracle/mcts_numpy.py exactly the way the reference wires a pgx
env + network into MCTS (core/evaluators/mcts/mcts.py:71-108,145-189; core/evaluators/alphazero.py:43-81;
core/common.py:32-103).  Written independently of the C/CUDA definition so the two can pin each other.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import mcts_numpy as M

f32 = np.float32
MASK32 = 0xFFFFFFFF


def mix32(x: int) -> int:
    x &= MASK32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & MASK32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & MASK32
    x ^= x >> 16
    return x


@dataclass
class SynthGame:
    F: int
    payload_bytes: int
    rho256: int
    tau1024: int
    max_depth: int
    seed: int

    @property
    def emb_row_bytes(self) -> List[int]:
        return [16] + ([self.payload_bytes] if self.payload_bytes > 0 else [])

    # --- scalar definitions -------------------------------------------------------------------------
    def init_h(self, env: int, episode: int) -> int:
        return mix32(self.seed * 0x9E3779B9 + env * 0x85EBCA6B + episode * 0xC2B2AE35 + 1)

    @staticmethod
    def step_h(h: int, a: int) -> int:
        return mix32(h * 0x9E3779B1 + a + 1)

    def legal(self, h: int, a: int) -> bool:
        return a == 0 or (mix32(h ^ (((a + 1) * 0x85EBCA6B) & MASK32)) & 0xFF) < self.rho256

    @staticmethod
    def logit(h: int, a: int) -> np.float32:
        u = mix32(h ^ ((a * 0x27D4EB2F + 0x9E3779B9) & MASK32)) >> 8
        return f32(f32(u) * f32(1.0 / 4194304.0) - f32(2.0))

    def terminal(self, h: int, depth: int) -> bool:
        return depth >= self.max_depth or (mix32(h ^ 0xC2B2AE35) & 0x3FF) < self.tau1024

    @staticmethod
    def reward(h: int) -> np.float32:
        return f32((mix32(h ^ 0x165667B1) % 3) - 1)

    @staticmethod
    def value(h: int) -> np.float32:
        return f32(f32(mix32(h ^ 0x27D4EB2F) >> 8) * f32(1.0 / 8388608.0) - f32(1.0))

    # --- embeddings ---------------------------------------------------------------------------------
    def make_emb(self, h: int, depth: int, player: int) -> List[np.ndarray]:
        core = np.array([h, depth, player, 0], dtype=np.uint32).astype(np.uint32).view(np.uint8).copy()
        out = [core]
        P = self.payload_bytes
        if P > 0:
            nw = (P + 3) // 4
            words = np.array([mix32(h ^ (((w + 1) * 0x9E3779B9) & MASK32)) for w in range(nw)], dtype="<u4")
            out.append(words.view(np.uint8)[:P].copy())
        return out

    @staticmethod
    def read_core(emb: List[np.ndarray]):
        c = np.asarray(emb[0], dtype=np.uint8).view("<u4")
        return int(c[0]), int(c[1]), int(c[2])

    def init_state(self, env: int, episode: int) -> List[np.ndarray]:
        return self.make_emb(self.init_h(env, episode), 0, 0)

    # --- the three plug-in functions ------------------------------------------------------------------
    def logits(self, h: int) -> np.ndarray:
        return np.array([self.logit(h, a) for a in range(self.F)], dtype=f32)

    def mask(self, h: int) -> np.ndarray:
        return np.array([self.legal(h, a) for a in range(self.F)], dtype=bool)

    def root_eval(self, emb, dir_noise: Optional[np.ndarray], dir_eps: float):
        """mcts.py:137-138 (dir_noise None) or alphazero.py:57-76."""
        h, _, _ = self.read_core(emb)
        pol = M.softmax(self.logits(h))
        val = self.value(h)
        if dir_noise is None:
            return pol, val
        noisy = ((f32(1.0 - dir_eps) * pol).astype(f32) + (f32(dir_eps) * np.asarray(dir_noise, f32)).astype(f32)).astype(f32)
        new_logits = M.tz_logf(np.maximum(noisy, M.FLT_MIN))
        masked = np.where(self.mask(h), new_logits, -M.FLT_MAX).astype(f32)
        return M.softmax(masked), val

    def leaf_eval(self, parent_emb, action: int):
        """env_step_fn + eval_fn + mcts.py:166-172.  Returns (new_emb, policy, value, terminated)."""
        h, depth, player = self.read_core(parent_emb)
        h2 = self.step_h(h, action)
        d2 = depth + 1
        term = self.terminal(h2, d2)
        masked = np.where(self.mask(h2), self.logits(h2), -M.FLT_MAX).astype(f32)
        pol = M.softmax(masked)
        val = self.reward(h2) if term else self.value(h2)
        return self.make_emb(h2, d2, 1 - player), pol, val, term


GAMES = {
    # name: (F, payload_bytes, rho256, tau1024, max_depth) -- shapes of BASELINE.json's pgx games, SURVEY.md 8d
    "tic_tac_toe": (9, 72, 154, 40, 9),
    "connect_four": (7, 272, 230, 12, 42),
    "othello": (65, 448, 38, 6, 60),
    "go_9x9": (82, 4080, 205, 2, 120),
    "2048": (4, 560, 218, 4, 200),
}


def make_game(name: str, seed: int) -> SynthGame:
    F, P, rho, tau, D = GAMES[name]
    return SynthGame(F=F, payload_bytes=P, rho256=rho, tau1024=tau, max_depth=D, seed=seed)


# ----------------------------------------------------------------------------------------------------
# per-tree drivers
# ----------------------------------------------------------------------------------------------------
def iterate(tree: M.Tree, game: SynthGame, cfg: M.SearchCfg, bp_noise=None) -> int:
    """MCTS.iterate mcts.py:145-189.  Returns the number of selector levels walked."""
    parent, action, levels = M.traverse(tree, cfg)
    parent_emb = [e[parent] for e in tree.emb]
    new_emb, pol, val, term = game.leaf_eval(parent_emb, action)
    M.expand(tree, parent, action, pol, val, term, new_emb, cfg)
    if cfg.weighted:
        M.weighted_backpropagate(tree, parent, cfg, bp_noise)
    else:
        M.backpropagate(tree, parent, val, cfg)
    return levels


def evaluate(tree: M.Tree, game: SynthGame, cfg: M.SearchCfg, root_emb, num_iterations: int,
             temperature: float, dir_noise=None, dir_eps: float = 0.25, root_noise=None, uniform01=None,
             bp_noise=None):
    """MCTS.evaluate mcts.py:71-108.  bp_noise: (S,F) or None.  Returns (action, policy_weights, visits)."""
    pol, val = game.root_eval(root_emb, dir_noise, dir_eps)
    M.set_root(tree, pol, val, root_emb)
    for s in range(num_iterations):
        iterate(tree, game, cfg, None if bp_noise is None else bp_noise[s])
    return M.root_action(tree, temperature, root_noise, uniform01)


def env_step(game: SynthGame, env: int, episode: int, emb, action: int):
    """core/common.py:82-99 for the synthetic game: returns (emb', episode', reset_flag)."""
    h, depth, player = game.read_core(emb)
    h2, d2 = game.step_h(h, action), depth + 1
    if game.terminal(h2, d2):
        return game.init_state(env, episode + 1), episode + 1, True
    return game.make_emb(h2, d2, 1 - player), episode, False
