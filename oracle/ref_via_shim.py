"""ORACLE / TEST INFRASTRUCTURE: runs the reference's UNMODIFIED source (/root/reference/core/...) one env at a time on
the NumPy emulation of the jax API in oracle/jaxshim (jax itself is not installable here; see oracle/jaxshim/README.md)
and returns results in the layout of tests/helpers.Result.  Used by tests/golden/make_golden.py to generate the
committed golden fixtures and by tests/test_golden.py to re-check them when /root/reference is present (it never is on
the GPU box, and nothing that runs there imports this module).

Driven through the reference's own entry points: `step_env_and_evaluator` (core/common.py:32-103) ->
`AlphaZero(MCTS | WeightedMCTS).evaluate` (core/evaluators/mcts/mcts.py:71-108, core/evaluators/alphazero.py:43-81) ->
`MCTS.step` / `reset` (mcts.py:387-414) -> `Tree.get_subtree` (core/trees/tree.py:220-269).
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import List

import numpy as np

from . import synth_numpy as SN

REFERENCE_ROOT = os.environ.get("TZ_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jaxshim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "trees", "tree.py"))


_mods = None


def load_reference():
    """Imports the reference's modules with the shim packages standing in for jax / chex / flax / optax / graphviz."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_ROOT}")
    for name in ("jax", "chex", "flax", "optax", "graphviz"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__shim__", False) and name == "jax":
            raise RuntimeError("a real `jax` is already imported in this process")
    sys.path.insert(0, SHIM_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import jax  # the shim

        assert getattr(jax, "__shim__", False)
        names = ["core.trees.tree", "core.evaluators.mcts.state", "core.evaluators.mcts.action_selection",
                 "core.evaluators.mcts.mcts", "core.evaluators.mcts.weighted_mcts", "core.evaluators.alphazero",
                 "core.common", "core.types"]
        _mods = {n.split(".")[-1]: importlib.import_module(n) for n in names}
        _mods["jax"] = jax
    finally:
        sys.path.remove(SHIM_ROOT)
        sys.path.remove(REFERENCE_ROOT)
    return _mods


def _tree_arrays(tree, payload_bytes: int) -> dict:
    d = tree.data
    N = tree.parents.shape[0]
    out = {
        "next_free_idx": np.int32(int(tree.next_free_idx)),
        "parents": np.asarray(tree.parents).astype(np.int32),
        "edge_map": np.asarray(tree.edge_map).astype(np.int32),
        "n": np.asarray(d.n).astype(np.int32),
        "p": np.asarray(d.p).astype(np.float32),
        "q": np.asarray(d.q).astype(np.float32),
        "terminated": np.asarray(d.terminated).astype(np.uint8),
    }
    if hasattr(d, "r"):
        out["r"] = np.asarray(d.r).astype(np.float32)
    out["emb0"] = np.ascontiguousarray(np.asarray(d.embedding["core"]).astype("<i4")).view(np.uint8).reshape(N, 16)
    if payload_bytes > 0:
        out["emb1"] = np.asarray(d.embedding["payload"]).astype(np.uint8)
    return out


def run_reference(s, snapshots: bool = False):
    """`s` is a tests/helpers.Schedule.  Returns (arrays, actions, pw, snapshots) with trees stacked over the batch."""
    if s.fma_backup:
        raise ValueError("fma_backup is a what-if about XLA, not reference source behaviour")
    R = load_reference()
    jax = R["jax"]
    jnp = jax.numpy
    StepMetadata = R["types"].StepMetadata
    g: SN.SynthGame = s.game
    F, P = g.F, g.payload_bytes

    def state(h, depth, player):
        emb = g.make_emb(h, depth, player)
        st = {"core": jnp.array(emb[0].view("<i4").copy(), dtype=jnp.int32)}
        if P > 0:
            st["payload"] = jnp.array(emb[1], dtype=jnp.uint8)
        return st

    def read(st):
        c = np.asarray(st["core"]).astype(np.int64)
        return int(c[0]) & 0xFFFFFFFF, int(c[1]), int(c[2])

    def metadata(h, depth, player, terminated):
        rew = g.reward(h)
        return StepMetadata(rewards=jnp.array(np.array([rew, rew], np.float32)), action_mask=jnp.array(g.mask(h)),
                            terminated=jnp.array(bool(terminated), dtype=jnp.bool_),
                            cur_player_id=jnp.array(player, dtype=jnp.int32), step=jnp.array(depth, dtype=jnp.int32))

    def eval_fn(env_state, params, key):  # core/types.py:31
        h, _, _ = read(env_state)
        return jnp.array(g.logits(h)), jnp.array(g.value(h), dtype=jnp.float32)

    def env_step_fn(env_state, action):  # core/types.py:27
        h, depth, player = read(env_state)
        h2, d2 = g.step_h(h, int(action)), depth + 1
        return state(h2, d2, 1 - player), metadata(h2, d2, 1 - player, g.terminal(h2, d2))

    base = R["weighted_mcts"].WeightedMCTS if s.weighted else R["mcts"].MCTS
    AS = R["action_selection"]
    # the `q_transform` constructor argument (action_selection.py:70,128): the reference's own normalize_q_values, or the
    # identity (TZ_QT_IDENTITY) as a plain Python callable -- both handed to the reference's UNMODIFIED selector classes
    qt4 = {0: AS.normalize_q_values, 1: (lambda q, n, parent_q, eps: q)}[getattr(s, "q_transform", 0)]
    if s.selector == 0:
        selector = AS.PUCTSelector(c=s.c, q_transform=qt4)
    else:
        # MuZeroPUCTSelector.__call__ passes FIVE positional arguments -- (discounted q, q, n, parent q, epsilon) -- to a
        # four-argument q_transform (action_selection.py:169), so it cannot run with its default.  The "arity fix" lives
        # HERE, not in the reference source: a five-argument adapter that drops the un-discounted q and forwards to the
        # four-argument function.  Fixtures made this way carry "arityfix" in their name.
        def qt5(discounted_q, _q, n, parent_q, eps):
            return qt4(discounted_q, n, parent_q, eps)

        selector = AS.MuZeroPUCTSelector(c1=s.c1, c2=s.c2, q_transform=qt5)
    kw = dict(eval_fn=eval_fn, action_selector=selector, branching_factor=F, max_nodes=s.N,
              num_iterations=s.S, discount=s.discount, temperature=s.temperature, tiebreak_noise=s.tiebreak_noise,
              persist_tree=s.persist_tree)
    if s.weighted:
        kw["q_temperature"] = s.q_temperature
    if s.dirichlet:
        ev = R["alphazero"].AlphaZero(base)(dirichlet_alpha=s.dir_alpha, dirichlet_epsilon=s.dir_eps, **kw)
    else:
        ev = base(**kw)

    cur = {"b": 0}

    class Tape(jax.random.Tape):
        # key paths: step key (m,) -> evaluate key (m,1) -> (m,1,0) [root sampling], (m,1,1,1) [dirichlet],
        # (m,1,0,i,1) [backprop noise of simulation i]; see core/common.py:71, mcts.py:94-103,168, alphazero.py:57
        def uniform(self, key, shape, minval, maxval):
            p, b = key.path, cur["b"]
            if len(p) == 3:
                return s.uniform01[p[0], b] if shape == () else s.root_noise[p[0], b]
            assert len(p) == 5 and p[4] == 1, p
            return s.bp_noise[p[0], p[3], b]

        def dirichlet(self, key, alpha):
            assert len(key.path) == 4
            return s.dir_noise[key.path[0], cur["b"]]

    jax.random.install_tape(Tape())
    actions = np.zeros((s.moves, s.B), np.int32)
    pw = np.zeros((s.moves, s.B, F), np.float32)
    finals: List[dict] = []
    snaps: List[List[dict]] = [[] for _ in range(s.moves)]
    for b in range(s.B):
        cur["b"] = b
        env_id, episode = b + s.env_offset, [0]
        env_state = state(g.init_h(env_id, 0), 0, 0)
        h0 = read(env_state)[0]
        md = metadata(h0, 0, 0, False)
        tree = ev.init(template_embedding=env_state)

        def env_init_fn(key):  # core/types.py:28
            episode[0] += 1
            h = g.init_h(env_id, episode[0])
            return state(h, 0, 0), metadata(h, 0, 0, False)

        for m in range(s.moves):
            if snapshots:
                orig_step, orig_reset = ev.step, ev.reset

                def rec_step(st, action, _m=m, _f=orig_step):
                    snaps[_m].append(_tree_arrays(st, P))
                    return _f(st, action)

                def rec_reset(st, _m=m, _f=orig_reset):
                    snaps[_m].append(_tree_arrays(st, P))
                    return _f(st)

                ev.step, ev.reset = rec_step, rec_reset
            out, env_state, md, _, _, _ = R["common"].step_env_and_evaluator(
                key=jax.random.Key((m,)), env_state=env_state, env_state_metadata=md, eval_state=tree, params=None,
                evaluator=ev, env_step_fn=env_step_fn, env_init_fn=env_init_fn, max_steps=1 << 30)
            if snapshots:
                del ev.step, ev.reset
            tree = out.eval_state
            actions[m, b] = int(out.action)
            pw[m, b] = np.asarray(out.policy_weights)
        finals.append(_tree_arrays(tree, P))
    stack = lambda lst: {k: np.stack([t[k] for t in lst]) for k in lst[0]}
    return stack(finals), actions, pw, ([stack(x) for x in snaps] if snapshots else None)


def run_reference_two_player(g: SN.SynthGame, ev_kw_1: dict, ev_kw_2: dict, p1_first, max_steps: int, dir_noise, root_noise,
                             uniform01, dir_alpha=0.3, dir_eps=0.25):
    """Two AlphaZero(MCTS) evaluators playing each other, one game at a time, through the reference's own
    `two_player_game_step` (core/common.py:146-232) driven exactly as `two_player_game` drives it (common.py:303-355: a
    turn for each player per scan step, skipped once `completed`), with who-moves-first given per game (`p1_first[b]`,
    the reference draws it with jax.random.randint, common.py:276).  Noise arrays are indexed [half_step, game].
    Returns per-half-step records: action (-1 where the game was already completed), p1 / p2 value estimates, completed,
    and the final outcomes (B, 2)."""
    R = load_reference()
    jax = R["jax"]
    jnp = jax.numpy
    StepMetadata = R["types"].StepMetadata
    C = R["common"]
    F, P = g.F, g.payload_bytes
    B = len(p1_first)

    def state(h, depth, player):
        emb = g.make_emb(h, depth, player)
        st = {"core": jnp.array(emb[0].view("<i4").copy(), dtype=jnp.int32)}
        if P > 0:
            st["payload"] = jnp.array(emb[1], dtype=jnp.uint8)
        return st

    def read(st):
        c = np.asarray(st["core"]).astype(np.int64)
        return int(c[0]) & 0xFFFFFFFF, int(c[1]), int(c[2])

    def metadata(h, depth, player, terminated):
        rew = g.reward(h)
        return StepMetadata(rewards=jnp.array(np.array([rew, rew], np.float32)), action_mask=jnp.array(g.mask(h)),  # as the stand-in's leaf: one reward for both seats
                            terminated=jnp.array(bool(terminated), dtype=jnp.bool_),
                            cur_player_id=jnp.array(player, dtype=jnp.int32), step=jnp.array(depth, dtype=jnp.int32))

    def eval_fn(env_state, params, key):
        h, _, _ = read(env_state)
        return jnp.array(g.logits(h)), jnp.array(g.value(h), dtype=jnp.float32)

    def env_step_fn(env_state, action):
        h, depth, player = read(env_state)
        h2, d2 = g.step_h(h, int(action)), depth + 1
        return state(h2, d2, 1 - player), metadata(h2, d2, 1 - player, g.terminal(h2, d2))

    def make_ev(kw):
        return R["alphazero"].AlphaZero(R["mcts"].MCTS)(
            dirichlet_alpha=dir_alpha, dirichlet_epsilon=dir_eps, eval_fn=eval_fn,
            action_selector=R["action_selection"].PUCTSelector(c=kw.get("c", 1.0)), branching_factor=F, max_nodes=kw["N"],
            num_iterations=kw["S"], discount=-1.0, temperature=kw.get("temperature", 1.0), tiebreak_noise=1e-8, persist_tree=True)

    ev1, ev2 = make_ev(ev_kw_1), make_ev(ev_kw_2)
    BASE = 1000
    cur = {"b": 0}

    class Tape(jax.random.Tape):
        # state.key = (BASE,) + (1,)*t at half-step t; step_key = state.key + (0,) (common.py:180); then as in run_reference:
        # evaluate key = step_key + (1,) (common.py:71); + (0,) root sampling (mcts.py:94-103); + (1, 1) dirichlet (alphazero.py:57)
        @staticmethod
        def _parse(p):
            assert p[0] == BASE
            t = 0
            while p[1 + t] == 1:
                t += 1
            assert p[1 + t] == 0
            return t, p[2 + t:]

        def uniform(self, key, shape, minval, maxval):
            t, rest = self._parse(key.path)
            assert rest == (1, 0), key.path
            return uniform01[t, cur["b"]] if shape == () else root_noise[t, cur["b"]]

        def dirichlet(self, key, alpha):
            t, rest = self._parse(key.path)
            assert rest == (1, 1, 1), key.path
            return dir_noise[t, cur["b"]]

    jax.random.install_tape(Tape())
    T = (max_steps // 2) * 2
    actions = np.full((T, B), -1, np.int32)
    p1v, p2v = np.zeros((T, B), np.float32), np.zeros((T, B), np.float32)
    completed = np.zeros((T, B), bool)
    outcomes = np.zeros((B, 2), np.float32)
    p1_nfi, p2_nfi = np.zeros((T, B), np.int32), np.zeros((T, B), np.int32)
    for b in range(B):
        cur["b"] = b
        h0 = g.init_h(b, 0)
        env_state, md = state(h0, 0, 0), metadata(h0, 0, 0, False)
        st = C.TwoPlayerGameState(
            key=jax.random.Key((BASE,)), env_state=env_state, env_state_metadata=md,
            p1_eval_state=ev1.init(template_embedding=env_state), p2_eval_state=ev2.init(template_embedding=env_state),
            p1_value_estimate=jnp.array(0.0, dtype=jnp.float32), p2_value_estimate=jnp.array(0.0, dtype=jnp.float32),
            outcomes=jnp.zeros((2,), dtype=jnp.float32), completed=jnp.zeros((), dtype=jnp.bool_))
        recorded = {}
        orig = C.step_env_and_evaluator

        def spy(**kw):  # records the action the active evaluator chose
            out = orig(**kw)
            recorded["action"] = int(out[0].action)
            return out

        C.step_env_and_evaluator = spy
        try:
            for t in range(T):
                use_p1 = bool(p1_first[b]) == (t % 2 == 0)
                if not bool(st.completed):  # common.py:305-317 / 327-339
                    # the key advances only when a step is taken; keep the tape's half-step index equal to t
                    st = st.replace(key=jax.random.Key((BASE,) + (1,) * t))
                    st = C.two_player_game_step(st, p1_evaluator=ev1, p2_evaluator=ev2, params=None, env_step_fn=env_step_fn,
                                                env_init_fn=None, use_p1=use_p1, max_steps=max_steps)
                    actions[t, b] = recorded["action"]
                p1v[t, b], p2v[t, b] = float(st.p1_value_estimate), float(st.p2_value_estimate)
                completed[t, b] = bool(st.completed)
                p1_nfi[t, b], p2_nfi[t, b] = int(st.p1_eval_state.next_free_idx), int(st.p2_eval_state.next_free_idx)
        finally:
            C.step_env_and_evaluator = orig
        outcomes[b] = np.asarray(st.outcomes)
    return dict(actions=actions, p1_value=p1v, p2_value=p2v, completed=completed, outcomes=outcomes, p1_nfi=p1_nfi, p2_nfi=p2_nfi)
