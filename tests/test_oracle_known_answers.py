"""Hand-derived known answers and reference edge cases, checked against BOTH oracles:
the literal NumPy restatement (oracle/mcts_numpy.py) and the C restatement (oracle/tz_oracle.c)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import mcts_numpy as M

f32 = np.float32


# ---- helpers: the same hand-built tree in both representations ---------------------------------------------------
def ka1_numpy():
    t = M.init_tree(6, 2, [4])
    t.parents[:] = [-1, 0, 0, 1, 2, 3]
    t.edge_map[:] = [[1, 2], [-1, 3], [4, -1], [5, -1], [-1, -1], [-1, -1]]
    t.next_free_idx = 6
    t.n[:] = [10, 11, 12, 13, 14, 15]
    t.q[:] = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]
    t.p[:] = np.arange(12, dtype=f32).reshape(6, 2) / 16
    t.terminated[:] = [0, 0, 0, 0, 1, 0]
    t.emb[0][:] = np.arange(24, dtype=np.uint8).reshape(6, 4) + 1
    return t


def to_host(trees, weighted=False):
    B, N, F = len(trees), trees[0].capacity, trees[0].branching_factor
    h = CO.HostTrees(B, N, F, [e.shape[1] for e in trees[0].emb], weighted=weighted)
    for b, t in enumerate(trees):
        h.next_free_idx[b] = t.next_free_idx
        h.parents[b], h.edge_map[b], h.n[b], h.p[b], h.q[b], h.terminated[b] = t.parents, t.edge_map, t.n, t.p, t.q, t.terminated
        for k, e in enumerate(t.emb):
            h.emb[k][b] = e
        if weighted:
            h.r[b] = t.r
    return h


def assert_host_equals(h, b, t):
    assert h.next_free_idx[b] == t.next_free_idx
    for name in ("parents", "edge_map", "n", "p", "q", "terminated"):
        assert np.array_equal(getattr(h, name)[b], getattr(t, name)), name
    for k, e in enumerate(t.emb):
        assert np.array_equal(h.emb[k][b], e)
    if t.r is not None:
        assert np.array_equal(h.r[b], t.r)


# ---- KA-1: get_subtree (tree.py:169-269) ------------------------------------------------------------------------
@pytest.mark.parametrize("action,exp_nfi,exp_parents,exp_edges,exp_rows", [
    (0, 3, [-1, 0, 1, -1, -1, -1], [[-1, 1], [2, -1], [-1, -1], [-1, -1], [-1, -1], [-1, -1]], [1, 3, 5]),
    (1, 2, [-1, 0, -1, -1, -1, -1], [[1, -1], [-1, -1], [-1, -1], [-1, -1], [-1, -1], [-1, -1]], [2, 4]),
])
def test_ka1_get_subtree(action, exp_nfi, exp_parents, exp_edges, exp_rows):
    old = ka1_numpy()
    t = old.copy()
    M.get_subtree(t, action)
    assert t.next_free_idx == exp_nfi
    assert t.parents.tolist() == exp_parents
    assert t.edge_map.tolist() == exp_edges
    for slot in range(6):
        if slot < exp_nfi:
            src = exp_rows[slot]
            assert t.n[slot] == old.n[src] and t.q[slot] == old.q[src] and t.terminated[slot] == old.terminated[src]
            assert np.array_equal(t.p[slot], old.p[src]) and np.array_equal(t.emb[0][slot], old.emb[0][src])
        else:  # erased data rows are zeros, not stale values (tree.py:236-238)
            assert t.n[slot] == 0 and t.q[slot] == 0 and t.terminated[slot] == 0
            assert not t.p[slot].any() and not t.emb[0][slot].any()
    h = to_host([old, old])
    CO.reroot(h, np.array([action, action], np.int32))
    assert_host_equals(h, 0, t)
    assert_host_equals(h, 1, t)


def test_ka1_labels_match_survey():
    t = ka1_numpy()
    old_idx, translation, erase = M._get_translation(t, 0)
    assert translation.tolist() == [-1, 0, -1, 1, -1, 2]
    assert erase.tolist() == [False, False, False, True, True, True]


# ---- KA-2: PUCT (action_selection.py:91-116) ---------------------------------------------------------------------
def ka2_tree():
    t = M.init_tree(4, 3, [])
    t.next_free_idx = 2
    t.parents[1] = 0
    t.edge_map[0, 0] = 1
    t.n[:2] = [3, 2]
    t.q[:2] = [f32(0.2), f32(-0.4)]
    t.p[0] = [0.5, 0.3, 0.2]
    return t


def test_ka2_puct_scores_and_action():
    t = ka2_tree()
    cfg = M.SearchCfg()
    q = (M.get_child_data(t, t.q, 0) * f32(-1.0)).astype(f32)
    n = M.get_child_data(t, t.n, 0)
    qn = M.normalize_q_values(q, n, t.q[0], 1e-8)
    assert qn.tolist() == [1.0, 0.0, 0.0]
    u = ((f32(1.0) * t.p[0]).astype(f32) * np.sqrt(f32(3))).astype(f32) / (n + 1).astype(f32)
    np.testing.assert_allclose(u, [0.28867513, 0.5196152, 0.34641016], rtol=1e-7)
    assert M.select_action(t, 0, cfg) == 0
    h = to_host([t])
    w = CO.HostWork(1, 3, [])
    CO.select(h, CO.make_cfg(), w)
    # child 0 exists and is not terminal -> descend; at the leaf-less child every score ties at 0 -> action 0
    assert w.parent[0] == 1 and w.action[0] == 0
    assert M.traverse(t, cfg)[:2] == (1, 0)


def test_ka2_tie_break_is_lowest_index_and_illegal_moves_score_zero():
    t = M.init_tree(4, 4, [])
    M.set_root(t, np.array([0.0, 0.5, 0.5, 0.0], f32), 0.3, [])
    assert M.select_action(t, 0, M.SearchCfg()) == 1  # p == 0 (illegal) scores 0; first of the tied maxima wins
    t.p[0] = 0
    assert M.select_action(t, 0, M.SearchCfg()) == 0  # all zero: argmax returns the first index


# ---- KA-3: backprop (mcts.py:231-262, 322) ----------------------------------------------------------------------
def test_ka3_backprop():
    t = M.init_tree(4, 3, [])
    t.next_free_idx = 2
    t.parents[1] = 0
    t.edge_map[0, 0] = 1
    t.n[:2] = [3, 2]
    t.q[:2] = [f32(0.2), f32(-0.4)]
    cfg = M.SearchCfg()
    h = to_host([t])
    pol = np.array([0.2, 0.3, 0.5], f32)
    M.expand(t, 1, 2, pol, f32(0.6), False, [], cfg)
    M.backpropagate(t, 1, f32(0.6), cfg)
    assert t.next_free_idx == 3 and t.edge_map[1, 2] == 2 and t.parents[2] == 1
    assert t.n[:3].tolist() == [4, 3, 1]
    assert t.q[2] == f32(0.6)
    assert t.q[1] == f32(f32(f32(f32(-0.4) * f32(2)) + f32(-0.6)) / f32(3))
    assert t.q[0] == f32(f32(f32(f32(0.2) * f32(3)) + f32(0.6)) / f32(4))
    np.testing.assert_allclose(t.q[:2], [0.3, -0.46666667], rtol=1e-6)
    w = CO.HostWork(1, 3, [])
    w.parent[0], w.action[0], w.policy[0], w.value[0], w.terminated[0] = 1, 2, pol, 0.6, 0
    CO.expand_backprop(h, CO.make_cfg(), w)
    assert_host_equals(h, 0, t)


# ---- edge cases the reference's semantics define (SURVEY.md "Semantics contract") --------------------------------
def test_full_tree_drops_the_node_but_still_backpropagates():  # Q7, tree.py:116-132, mcts.py:35-36
    t = M.init_tree(2, 2, [3])
    M.set_root(t, np.array([0.5, 0.5], f32), 0.1, [np.array([1, 2, 3], np.uint8)])
    cfg = M.SearchCfg()
    M.expand(t, 0, 0, np.array([1, 0], f32), f32(0.5), False, [np.array([4, 5, 6], np.uint8)], cfg)
    M.backpropagate(t, 0, f32(0.5), cfg)
    assert t.next_free_idx == 2
    before = t.copy()
    h = to_host([t])
    M.expand(t, 0, 1, np.array([0, 1], f32), f32(-0.25), True, [np.array([7, 8, 9], np.uint8)], cfg)
    M.backpropagate(t, 0, f32(-0.25), cfg)
    assert t.next_free_idx == 2 and t.edge_map[0, 1] == -1
    assert np.array_equal(t.p, before.p) and np.array_equal(t.emb[0], before.emb[0]) and np.array_equal(t.terminated, before.terminated)
    assert t.n[0] == before.n[0] + 1 and t.q[0] != before.q[0]
    w = CO.HostWork(1, 2, [3])
    w.parent[0], w.action[0], w.policy[0], w.value[0], w.terminated[0] = 0, 1, [0, 1], -0.25, 1
    w.emb_new[0][0] = [7, 8, 9]
    CO.expand_backprop(h, CO.make_cfg(), w)
    assert_host_equals(h, 0, t)


def test_terminal_child_is_revisited_and_overwritten():  # Q6, mcts.py:179,208-211
    t = M.init_tree(4, 2, [2])
    M.set_root(t, np.array([1.0, 0.0], f32), 0.0, [np.array([1, 1], np.uint8)])
    cfg = M.SearchCfg()
    M.expand(t, 0, 0, np.array([0.5, 0.5], f32), f32(1.0), True, [np.array([2, 2], np.uint8)], cfg)
    M.backpropagate(t, 0, f32(1.0), cfg)
    assert M.traverse(t, cfg)[:2] == (0, 0)  # stops AT the root: the chosen child exists but is terminal
    h = to_host([t])
    M.expand(t, 0, 0, np.array([0.25, 0.75], f32), f32(-1.0), True, [np.array([3, 3], np.uint8)], cfg)
    assert t.next_free_idx == 2 and t.n[1] == 2 and t.q[1] == f32(0.0)
    assert t.p[1].tolist() == [0.25, 0.75] and t.emb[0][1].tolist() == [3, 3]
    w = CO.HostWork(1, 2, [2])
    CO.select(h, CO.make_cfg(), w)
    assert (w.parent[0], w.action[0]) == (0, 0)
    w.policy[0], w.value[0], w.terminated[0] = [0.25, 0.75], -1.0, 1
    w.emb_new[0][0] = [3, 3]
    hb = to_host([t])  # numpy state after expand but before backprop
    M.backpropagate(t, 0, f32(-1.0), cfg)
    CO.expand_backprop(h, CO.make_cfg(), w)
    assert_host_equals(h, 0, t)
    del hb


def test_set_root_keeps_persisted_statistics():  # Q8, mcts.py:376-384, tree.py:147
    t = M.init_tree(3, 2, [1])
    M.set_root(t, np.array([0.5, 0.5], f32), 0.7, [np.array([9], np.uint8)])
    assert (t.n[0], t.q[0], t.next_free_idx) == (1, f32(0.7), 1)
    t.n[0], t.q[0], t.terminated[0] = 5, f32(-0.1), 1
    M.set_root(t, np.array([0.1, 0.9], f32), 0.3, [np.array([8], np.uint8)])
    assert (t.n[0], t.q[0], t.terminated[0]) == (5, f32(-0.1), 1)
    assert t.p[0].tolist() == [f32(0.1), f32(0.9)] and t.emb[0][0, 0] == 8


def test_step_on_absent_child_erases_everything():  # Q14, tree.py:201-203
    t = ka1_numpy()
    t.edge_map[0, 1] = -1
    h = to_host([t])
    M.get_subtree(t, 1)
    assert t.next_free_idx == 0 and (t.parents == -1).all() and (t.edge_map == -1).all()
    assert not t.n.any() and not t.p.any() and not t.emb[0].any()
    CO.reroot(h, np.array([1], np.int32))
    assert_host_equals(h, 0, t)


def test_reset_flag_semantics_of_the_fused_reroot():  # common.py:89-94
    t = ka1_numpy()
    h = to_host([t, t, t])
    CO.reroot(h, np.array([0, 0, 0], np.int32), np.array([0, 1, 2], np.uint8))
    a = t.copy()
    M.get_subtree(a, 0)
    assert_host_equals(h, 0, a)  # flag 0: step
    b = t.copy()
    M.reset(b)
    assert_host_equals(h, 1, b)  # flag 1: reset
    assert_host_equals(h, 2, t)  # flag 2: untouched


def test_root_action_paths():  # Q12, mcts.py:265-296
    t = M.init_tree(4, 3, [])
    M.set_root(t, np.array([0.2, 0.3, 0.5], f32), 0.0, [])
    a, pw, vis = M.root_action(t, 0.0, noise=np.array([0, 0, 1e-7], f32))
    assert pw.tolist() == [f32(1 / 3)] * 3 and vis.tolist() == [0, 0, 0] and a == 2  # no visits: uniform, noise decides
    # the reference's default tiebreak_noise (1e-8) is below fp32 resolution at 1/3: absorbed, first index wins
    assert M.root_action(t, 0.0, noise=np.array([0, 0, 1e-9], f32))[0] == 0
    for act, n in ((0, 3), (2, 1)):
        M.expand(t, 0, act, np.array([1, 0, 0], f32), f32(0), False, [], M.SearchCfg())
        t.n[t.edge_map[0, act]] = n
    a, pw, vis = M.root_action(t, 0.0, noise=np.zeros(3, f32))
    assert vis.tolist() == [3, 0, 1] and pw.tolist() == [0.75, 0.0, 0.25] and a == 0
    # temperature 1: jax.random.choice = searchsorted(cumsum(p), total * (1 - u))
    assert M.root_action(t, 1.0, uniform01=f32(0.9))[0] == 0  # r = 0.1
    assert M.root_action(t, 1.0, uniform01=f32(0.2))[0] == 2  # r = 0.8 > 0.75
    h = to_host([t, t])
    act, pw2, vis2, q0 = CO.root_action(h, 1.0, None, np.array([0.9, 0.2], f32))
    assert act.tolist() == [0, 2] and np.array_equal(pw2[0], pw) and np.array_equal(vis2[1], vis)


def test_weighted_backprop_node_without_visited_children():  # Q15, weighted_mcts.py:102-146
    t = M.init_tree(4, 3, [], weighted=True)
    M.set_root(t, np.array([0.2, 0.3, 0.5], f32), 0.4, [])
    assert t.r[0] == f32(0.4)
    cfg = M.SearchCfg(weighted=True)
    M.weighted_backpropagate(t, 0, cfg)
    # uniform weights over all-zero values -> q_w = 0 ; q' = (0 * 1 + r) / 2
    assert t.n[0] == 2 and t.q[0] == f32(0.2)
