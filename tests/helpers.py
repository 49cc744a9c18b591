"""Shared drivers for the parity tests: the same self-play schedule run through
 (a) the literal NumPy restatement (oracle/mcts_numpy.py, per tree),
 (b) the C oracle step by step in the device's batched call order (oracle/tz_oracle.c via oracle/c_oracle.py),
 (c) the C oracle tree-major (tzo_selfplay), and -- in the gpu tests -- (d) the CUDA path through the C-ABI.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

from oracle import c_oracle as CO
from oracle import mcts_numpy as M
from oracle import synth_numpy as SN


@dataclass
class Schedule:
    """Everything random is drawn up front so every implementation consumes identical inputs."""
    game: SN.SynthGame
    B: int
    N: int
    S: int
    moves: int
    temperature: float = 1.0
    persist_tree: bool = True
    dirichlet: bool = True
    dir_eps: float = 0.25
    dir_alpha: float = 0.3
    weighted: bool = False
    q_temperature: float = 1.0
    selector: int = 0
    c1: float = 1.25  # MuZeroPUCTSelector (selector == 1)
    c2: float = 19652.0
    q_transform: int = 0  # include/tz_abi.h TZ_QT_*
    discount: float = -1.0
    c: float = 1.0
    fma_backup: bool = False
    programmatic: bool = False  # TzSearchCfg.programmatic (CUDA path only: launch mechanics, not semantics)
    sim_warps: int = 0  # TzSearchCfg.sim_warps (CUDA path only: warps per tree, not semantics)
    tiebreak_noise: float = 1e-8
    seed: int = 0
    env_offset: int = 0

    def __post_init__(self):
        rng = np.random.default_rng(self.seed + 7919 * self.env_offset)
        B, F, mv = self.B, self.game.F, self.moves
        self.dir_noise = (rng.dirichlet([self.dir_alpha] * F, size=(mv, B)).astype(np.float32)
                          if self.dirichlet else None)
        self.root_noise = (rng.random((mv, B, F), dtype=np.float32) * np.float32(self.tiebreak_noise)).astype(np.float32)
        self.uniform01 = rng.random((mv, B), dtype=np.float32)
        self.bp_noise = ((rng.random((mv, self.S, B, F), dtype=np.float32) * np.float32(self.tiebreak_noise)).astype(np.float32)
                         if (self.weighted and self.q_temperature == 0) else None)

    def np_cfg(self) -> M.SearchCfg:
        return M.SearchCfg(selector=self.selector, c=self.c, c1=self.c1, c2=self.c2, discount=self.discount, weighted=self.weighted,
                           q_temperature=self.q_temperature, fma_backup=self.fma_backup, q_transform=self.q_transform)

    def c_cfg(self):
        return CO.make_cfg(selector=self.selector, c=self.c, c1=self.c1, c2=self.c2, discount=self.discount, weighted=self.weighted,
                           q_temperature=self.q_temperature, fma_backup=self.fma_backup, q_transform=self.q_transform)

    def c_game(self):
        g = self.game
        return CO.make_game(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)


@dataclass
class Result:
    arrays: dict  # final tree arrays, batched
    actions: np.ndarray  # [moves,B]
    pw: np.ndarray  # [moves,B,F]
    snapshots: Optional[List[dict]] = None  # per move, after search (before re-root)


def run_numpy(s: Schedule, snapshots: bool = False) -> Result:
    g, cfg = s.game, s.np_cfg()
    rbs = g.emb_row_bytes
    trees = [M.init_tree(s.N, g.F, rbs, weighted=s.weighted) for _ in range(s.B)]
    actions = np.zeros((s.moves, s.B), np.int32)
    pw = np.zeros((s.moves, s.B, g.F), np.float32)
    snaps = [] if snapshots else None
    for b in range(s.B):
        env_id = b + s.env_offset
        episode = 0
        emb = g.init_state(env_id, episode)
        for m in range(s.moves):
            a, w, _ = SN.evaluate(
                trees[b], g, cfg, emb, s.S, s.temperature,
                dir_noise=None if s.dir_noise is None else s.dir_noise[m, b], dir_eps=s.dir_eps,
                root_noise=s.root_noise[m, b], uniform01=s.uniform01[m, b],
                bp_noise=None if s.bp_noise is None else s.bp_noise[m, :, b])
            actions[m, b], pw[m, b] = a, w
            if snapshots:
                if b == 0:
                    snaps.append([])
                snaps[m].append(trees[b].copy())
            emb, episode, rf = SN.env_step(g, env_id, episode, emb, a)
            if rf:
                M.reset(trees[b])
            else:
                M.step(trees[b], a, s.persist_tree)
    res = Result(stack_trees(trees), actions, pw)
    if snapshots:
        res.snapshots = [stack_trees(ts) for ts in snaps]
    return res


def stack_trees(trees: List[M.Tree]) -> dict:
    d = {
        "next_free_idx": np.array([t.next_free_idx for t in trees], np.int32),
        "parents": np.stack([t.parents for t in trees]),
        "edge_map": np.stack([t.edge_map for t in trees]),
        "n": np.stack([t.n for t in trees]),
        "p": np.stack([t.p for t in trees]),
        "q": np.stack([t.q for t in trees]),
        "terminated": np.stack([t.terminated for t in trees]),
    }
    if trees[0].r is not None:
        d["r"] = np.stack([t.r for t in trees])
    for k in range(len(trees[0].emb)):
        d[f"emb{k}"] = np.stack([t.emb[k] for t in trees])
    return d


def run_c_stepwise(s: Schedule, snapshots: bool = False) -> Result:
    """The C oracle driven in the batched launch order the device uses."""
    g, cg, cfg = s.game, s.c_game(), s.c_cfg()
    B, F, P = s.B, g.F, g.payload_bytes
    rbs = g.emb_row_bytes
    t = CO.HostTrees(B, s.N, F, rbs, weighted=s.weighted)
    w = CO.HostWork(B, F, rbs, with_noise=s.bp_noise is not None)
    episode = np.zeros((B,), np.int32)
    core = np.zeros((B, 4), np.int32)
    payload = np.zeros((B, P), np.uint8) if P > 0 else None
    CO.synth_init_states(cg, B, s.env_offset, episode, core, payload)
    rp = np.zeros((B, F), np.float32)
    rv = np.zeros((B,), np.float32)
    reset_flag = np.zeros((B,), np.uint8)
    actions = np.zeros((s.moves, B), np.int32)
    pw = np.zeros((s.moves, B, F), np.float32)
    snaps = [] if snapshots else None
    for m in range(s.moves):
        CO.synth_root(cg, B, core, None if s.dir_noise is None else np.ascontiguousarray(s.dir_noise[m]), s.dir_eps, rp, rv)
        CO.set_root(t, rp, rv, [core.view(np.uint8).reshape(B, 16)] + ([payload] if P > 0 else []))
        for it in range(s.S):
            CO.select(t, cfg, w)
            CO.synth_leaf(cg, B, w.emb_parent[0].view(np.int32).reshape(B, 4), w.action, w.policy, w.value, w.terminated,
                          w.emb_new[0].view(np.int32).reshape(B, 4), w.emb_new[1] if P > 0 else None)
            if s.bp_noise is not None:
                w.backprop_noise[:] = s.bp_noise[m, it]
            CO.expand_backprop(t, cfg, w)
        act, pwm, _, _ = CO.root_action(t, s.temperature, np.ascontiguousarray(s.root_noise[m]),
                                        np.ascontiguousarray(s.uniform01[m]))
        actions[m], pw[m] = act, pwm
        if snapshots:
            snaps.append({k: v.copy() for k, v in t.arrays().items()})
        CO.synth_env_step(cg, B, s.env_offset, act, core, payload, episode, reset_flag)
        CO.reroot(t, act, reset_flag, s.persist_tree)
    res = Result({k: v.copy() for k, v in t.arrays().items()}, actions, pw, snaps)
    res.stats = t.stats.copy()
    return res


def run_c_treemajor(s: Schedule, nthreads: int = 0) -> Result:
    g, cg, cfg = s.game, s.c_game(), s.c_cfg()
    B, F, P = s.B, g.F, g.payload_bytes
    t = CO.HostTrees(B, s.N, F, g.emb_row_bytes, weighted=s.weighted)
    episode = np.zeros((B,), np.int32)
    core = np.zeros((B, 4), np.int32)
    payload = np.zeros((B, P), np.uint8) if P > 0 else None
    CO.synth_init_states(cg, B, s.env_offset, episode, core, payload)
    actions, pw = CO.selfplay(t, cfg, cg, s.S, s.moves, s.temperature, s.persist_tree, s.env_offset, s.dir_noise, s.dir_eps,
                              s.root_noise, s.uniform01, core, payload, episode, nthreads)
    res = Result({k: v.copy() for k, v in t.arrays().items()}, actions, pw)
    res.stats = t.stats.copy()
    return res


def assert_trees_equal(a: dict, b: dict, what: str = ""):
    """Bit-exact on integers / bytes; floats compared with == (so -0.0 == +0.0, NaN never expected)."""
    assert set(a) == set(b), (what, set(a) ^ set(b))
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and x.dtype == y.dtype, (what, k, x.shape, y.shape, x.dtype, y.dtype)
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            raise AssertionError(f"{what}: '{k}' differs at {len(bad)} places, first {bad[0].tolist()}: "
                                 f"{x[tuple(bad[0])]!r} vs {y[tuple(bad[0])]!r}")


def check_invariants(arr: dict):
    """Structural invariants every tree must satisfy."""
    nfi, parents, edge = arr["next_free_idx"], arr["parents"], arr["edge_map"]
    B, N = parents.shape
    for b in range(B):
        k = int(nfi[b])
        assert 0 <= k <= N
        assert np.all(parents[b, k:] == -1) and np.all(edge[b, k:] == -1)
        for name, v in arr.items():
            if name in ("next_free_idx", "parents", "edge_map", "stats"):
                continue
            assert not np.any(v[b, k:]), (name, b)
        if k == 0:
            continue
        assert parents[b, 0] == -1
        idx = np.arange(1, k)
        assert np.all(parents[b, 1:k] < idx) and np.all(parents[b, 1:k] >= 0)
        e = edge[b, :k]
        kids = e[e >= 0]
        assert len(np.unique(kids)) == len(kids) == k - 1  # every non-root node is the child of exactly one edge
        for i in range(1, k):
            assert np.sum(e[parents[b, i]] == i) == 1


# ----------------------------------------------------------------------------------------------------
# CUDA path (only imported / called by the gpu tests)
# ----------------------------------------------------------------------------------------------------
def tree_to_numpy(tree) -> dict:
    d = tree.data
    out = {
        "next_free_idx": tree.next_free_idx.cpu().numpy(),
        "parents": tree.parents.cpu().numpy(),
        "edge_map": tree.edge_map.cpu().numpy(),
        "n": d.n.cpu().numpy(),
        "p": d.p.cpu().numpy(),
        "q": d.q.cpu().numpy(),
        "terminated": d.terminated.cpu().numpy().astype(np.uint8),
    }
    if d.r is not None:
        out["r"] = d.r.cpu().numpy()
    for k, leaf in enumerate(tree._emb_leaves):
        B, N = leaf.shape[:2]
        out[f"emb{k}"] = leaf.contiguous().view(torch.uint8).reshape(B, N, -1).cpu().numpy() if leaf.dtype != torch.uint8 \
            else leaf.reshape(B, N, -1).cpu().numpy()
    return out


def assert_child_stats_consistent(tree, what=""):
    """The derived child_stats table the kernels maintain incrementally equals a from-scratch rebuild."""
    kept = tree.child_stats.clone()
    tree.rebuild_child_stats()
    assert torch.equal(kept, tree.child_stats), f"{what}: child_stats drifted from edge_map / q / n / terminated"


def assert_best_table_consistent(tree, ev, what="", expect_known=True):
    """Every known entry of the derived best-table (the selector's cached decision per node) equals the selector
    re-evaluated on the node's current rows; rows past next_free_idx hold the null entry."""
    import ctypes as C

    from turbozero_b200 import _abi

    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    cfg = ev._cfg()
    _abi.check(_abi.lib().tz_selftest_best(C.byref(tree.struct()), C.byref(cfg), out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "tz_selftest_best")
    bad, known = (int(x) for x in out.tolist())
    assert bad == 0, f"{what}: {bad} stale best-table entries (of {known} known)"
    if expect_known:
        assert known > 0, f"{what}: the best-table is empty"


def make_cuda_evaluator(s: Schedule, game):
    import turbozero_b200 as tz
    from standin.synthetic import make_synthetic_evaluator

    qt = {0: tz.normalize_q_values, 1: "identity"}[s.q_transform]
    sel = tz.PUCTSelector(c=s.c, q_transform=qt) if s.selector == 0 else tz.MuZeroPUCTSelector(c1=s.c1, c2=s.c2, q_transform=qt)
    base = tz.WeightedMCTS if s.weighted else tz.MCTS
    kw = dict(action_selector=sel, max_nodes=s.N, num_iterations=s.S, discount=s.discount, temperature=s.temperature,
              tiebreak_noise=s.tiebreak_noise, persist_tree=s.persist_tree)
    if s.weighted:
        kw["q_temperature"] = s.q_temperature
    ev = make_synthetic_evaluator(base, game, dir_eps=s.dir_eps, fma_backup=s.fma_backup, programmatic=s.programmatic, **kw)
    ev.sim_warps = s.sim_warps
    return ev


def run_cuda_api(s: Schedule, fused: bool = True, snapshots: bool = False) -> Result:
    """The CUDA path through the Python mirror of the reference API (MCTS.evaluate / iterate / step)."""
    import torch
    from standin.synthetic import SyntheticGame

    g = s.game
    game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    ev = make_cuda_evaluator(s, game)
    B, F = s.B, g.F
    tree = ev.init_batched(B, game.template_embedding(), stats=True)
    state, episode = game.init_states(B, s.env_offset)
    reset_flag = torch.zeros((B,), dtype=torch.uint8, device="cuda")
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    actions = np.zeros((s.moves, B), np.int32)
    pw = np.zeros((s.moves, B, F), np.float32)
    snaps = [] if snapshots else None
    for m in range(s.moves):
        dn = None if s.dir_noise is None else dev(s.dir_noise[m])
        rn, u01 = dev(s.root_noise[m]), dev(s.uniform01[m])
        bpn = None if s.bp_noise is None else dev(s.bp_noise[m])
        if fused:
            out = ev.evaluate(None, tree, state, None, None, None, leaf_fn=game.leaf_fn, root_noise=rn, uniform01=u01,
                              backprop_noise=bpn, dirichlet_noise=dn)
            act, pwm = out.action, out.policy_weights
        else:
            ev.update_root(None, tree, state, None, dirichlet_noise=dn)
            for it in range(s.S):
                ev.iterate(None, tree, None, None, leaf_fn=game.leaf_fn, backprop_noise=None if bpn is None else bpn[it])
            act, pwm = ev.sample_root_action(None, tree, root_noise=rn, uniform01=u01)
        actions[m], pw[m] = act.cpu().numpy(), pwm.cpu().numpy()
        if snapshots:
            snaps.append(tree_to_numpy(tree))
        assert_best_table_consistent(tree, ev, f"after the search of move {m}", expect_known=s.S > 0)
        assert_child_stats_consistent(tree, f"after the search of move {m}")
        game.env_step(state, act, episode, reset_flag, s.env_offset)
        ev.step(tree, act, reset_mask=reset_flag)
        assert_best_table_consistent(tree, ev, f"after re-rooting for move {m}", expect_known=False)
        assert_child_stats_consistent(tree, f"after re-rooting for move {m}")
    res = Result(tree_to_numpy(tree), actions, pw, snaps)
    res.stats = tree.stats.cpu().numpy().astype(np.uint64)
    return res


def run_cuda_selfplay(s: Schedule, use_path: bool = True, graph: bool = False, pipelines: int = 1, use_spill: bool = True) -> Result:
    """The CUDA path with the whole simulation loop inside the C-ABI (tz_search + tz_synth_leaf_cb)."""
    import torch
    from standin.synthetic import SyntheticGame, SyntheticSelfPlay

    g = s.game
    game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    ev = make_cuda_evaluator(s, game)
    sp = SyntheticSelfPlay(game, ev, s.B, env_offset=s.env_offset, dirichlet=s.dirichlet, use_path=use_path,
                           pipelines=pipelines, use_spill=use_spill)
    sp.dir_eps = s.dir_eps
    actions = np.zeros((s.moves, s.B), np.int32)
    pw = np.zeros((s.moves, s.B, g.F), np.float32)
    cg = None
    for m in range(s.moves):
        if s.dir_noise is not None:
            sp.dir_noise.copy_(torch.from_numpy(s.dir_noise[m]))
        sp.root_noise.copy_(torch.from_numpy(s.root_noise[m]))
        sp.uniform01.copy_(torch.from_numpy(s.uniform01[m]))
        if graph:
            if cg is None:
                # capture one move; capture does not execute, so replay it for move 0 as well
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(cg, stream=side):
                        sp.move()
                torch.cuda.current_stream().wait_stream(side)
            cg.replay()
        else:
            sp.move()
        actions[m], pw[m] = sp.action.cpu().numpy(), sp.policy_weights.cpu().numpy()
    assert_best_table_consistent(sp.tree, ev, "after self-play", expect_known=False)
    assert_child_stats_consistent(sp.tree, "after self-play")
    res = Result(tree_to_numpy(sp.tree), actions, pw)
    res.stats = sp.tree.stats.cpu().numpy().astype(np.uint64)
    return res
