"""ORACLE (test infrastructure, not product code) -- literal NumPy restatement of the reference MCTS path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (turbozero_b200/) never does.

PARITY PINNING: the reference (lowrollr/turbozero) ships no tests, golden vectors or fixtures, and JAX is not
installable here.  This restatement is pinned by (i) golden fixtures produced by executing the reference's UNMODIFIED
source files on a NumPy emulation of the jax API (oracle/jaxshim, oracle/ref_via_shim.py, tests/golden/make_golden.py;
checked by tests/test_golden.py), (ii) hand-derived known answers (SURVEY.md section 8c,
tests/test_oracle_known_answers.py) and (iii) agreement with an independently written C restatement
(oracle/tz_oracle.c).  It is NOT pinned against XLA's floating-point code generation (FMA contraction, XLA's
exp/log/pow): "parity unpinned" still applies to those bits, see DESIGN.md "Residual risk".  It is written one tree at
a time, which is exactly what the reference code expresses before `jax.vmap` (core/training/train.py:613) batches it.

Every function cites the reference lines it restates (paths relative to the reference repo root).
float32 everywhere; every float op is a single individually rounded IEEE operation, float sums use the
path's canonical order (`canon_sum`), exp/log/pow are the path's own definitions (include/tz_math.h).
JAX indexing rules relied on by the reference and restated explicitly here: negative indices wrap,
out-of-bounds scatter updates are dropped.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

f32 = np.float32
i32 = np.int32
NULL_INDEX = -1  # core/trees/tree.py:21
ROOT_INDEX = 0  # core/trees/tree.py:23
FLT_MAX = np.finfo(np.float32).max
FLT_MIN = np.finfo(np.float32).tiny
FLT_EPS = np.finfo(np.float32).eps

SEL_PUCT = 0
SEL_MUZERO = 1


# ----------------------------------------------------------------------------------------------------
# deterministic math (include/tz_math.h restated with numpy float32 scalars/arrays)
# ----------------------------------------------------------------------------------------------------
def tz_expf(x):
    x = np.asarray(x, dtype=f32)
    xc = np.minimum(x, f32(88.0))
    xc = np.where(x < f32(-86.0), f32(0.0), xc).astype(f32)
    k = np.rint(xc * f32(1.44269504088896341)).astype(f32)
    r = (xc - k * f32(0.693359375)).astype(f32)
    r = (r - k * f32(-2.12194440e-4)).astype(f32)
    p = np.full_like(r, f32(1.9875691500e-4))
    for c in (1.3981999507e-3, 8.3334519073e-3, 4.1665795894e-2, 1.6666665459e-1, 5.0000001201e-1):
        p = (p * r).astype(f32) + f32(c)
    r2 = (r * r).astype(f32)
    y = (((p * r2).astype(f32) + r).astype(f32) + f32(1.0)).astype(f32)
    ki = k.astype(np.int64)
    k1 = np.trunc(ki / 2).astype(np.int64)  # C integer division truncates toward zero
    k2 = ki - k1
    s1 = np.ldexp(f32(1.0), k1).astype(f32)
    s2 = np.ldexp(f32(1.0), k2).astype(f32)
    out = ((y * s1).astype(f32) * s2).astype(f32)
    return np.where(x < f32(-86.0), f32(0.0), out).astype(f32)


def tz_logf(x):
    x = np.maximum(np.asarray(x, dtype=f32), FLT_MIN).astype(f32)
    u = x.view(np.uint32) if x.ndim else np.array(x, dtype=f32).reshape(1).view(np.uint32)
    u = u.reshape(x.shape)
    e = ((u >> np.uint32(23)) & np.uint32(0xFF)).astype(np.int64) - 126
    m = ((u & np.uint32(0x007FFFFF)) | np.uint32(0x3F000000)).view(f32)
    small = m < f32(0.707106781186547524)
    e = np.where(small, e - 1, e)
    m = np.where(small, ((m + m).astype(f32) - f32(1.0)).astype(f32), (m - f32(1.0)).astype(f32)).astype(f32)
    z = (m * m).astype(f32)
    y = np.full_like(m, f32(7.0376836292e-2))
    for c in (-1.1514610310e-1, 1.1676998740e-1, -1.2420140846e-1, 1.4249322787e-1,
              -1.6668057665e-1, 2.0000714765e-1, -2.4999993993e-1, 3.3333331174e-1):
        y = ((y * m).astype(f32) + f32(c)).astype(f32)
    y = ((y * m).astype(f32) * z).astype(f32)
    fe = e.astype(f32)
    y = (y + (f32(-2.12194440e-4) * fe).astype(f32)).astype(f32)
    y = (y + (f32(-0.5) * z).astype(f32)).astype(f32)
    r = (m + y).astype(f32)
    r = (r + (f32(0.693359375) * fe).astype(f32)).astype(f32)
    return r


def tz_powf(x, y):
    x = np.asarray(x, dtype=f32)
    y = f32(y)
    if y == f32(1.0):
        return x.copy()
    out = tz_expf((y * tz_logf(x)).astype(f32))
    return np.where(x < FLT_MIN, f32(0.0), out).astype(f32)


def canon_sum(x) -> np.float32:
    """The path's canonical float32 sum order: 32 strided partial sums (element j goes to slot j % 32,
    accumulated in increasing j), then an xor-butterfly with offsets 16, 8, 4, 2, 1."""
    x = np.asarray(x, dtype=f32).ravel()
    v = np.zeros(32, dtype=f32)
    for j0 in range(0, x.size, 32):
        chunk = x[j0:j0 + 32]
        v[:chunk.size] = (v[:chunk.size] + chunk).astype(f32)
    lanes = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        v = (v + v[lanes ^ off]).astype(f32)
    return f32(v[0])


def softmax(logits) -> np.ndarray:
    """jax.nn.softmax: exp(x - max) / sum(exp(x - max)) with this path's exp and sum."""
    logits = np.asarray(logits, dtype=f32)
    m = logits.max()
    e = tz_expf((logits - m).astype(f32))
    return (e / canon_sum(e)).astype(f32)


# ----------------------------------------------------------------------------------------------------
# Tree (core/trees/tree.py) holding MCTSNode / WeightedMCTSNode (state.py:12-30, weighted_mcts.py:14-17)
# ----------------------------------------------------------------------------------------------------
@dataclass
class Tree:
    next_free_idx: int
    parents: np.ndarray  # (N,) int32
    edge_map: np.ndarray  # (N,F) int32
    n: np.ndarray  # (N,) int32
    p: np.ndarray  # (N,F) float32
    q: np.ndarray  # (N,) float32
    terminated: np.ndarray  # (N,) uint8
    emb: List[np.ndarray] = field(default_factory=list)  # each (N,row_k) uint8 (opaque bytes)
    r: Optional[np.ndarray] = None  # (N,) float32, weighted only

    @property
    def capacity(self) -> int:  # tree.py:26-28
        return self.parents.shape[0]

    @property
    def branching_factor(self) -> int:  # tree.py:32-34
        return self.edge_map.shape[1]

    def copy(self) -> "Tree":
        return Tree(int(self.next_free_idx), self.parents.copy(), self.edge_map.copy(), self.n.copy(),
                    self.p.copy(), self.q.copy(), self.terminated.copy(), [e.copy() for e in self.emb],
                    None if self.r is None else self.r.copy())


def init_tree(max_nodes: int, branching_factor: int, emb_row_bytes: List[int], weighted: bool = False) -> Tree:
    """tree.py:281-298 + mcts.py:417-432: indices -1, every data leaf zero, next_free_idx 0."""
    N, F = max_nodes, branching_factor
    return Tree(
        next_free_idx=0,
        parents=np.full((N,), NULL_INDEX, dtype=i32),
        edge_map=np.full((N, F), NULL_INDEX, dtype=i32),
        n=np.zeros((N,), dtype=i32),
        p=np.zeros((N, F), dtype=f32),
        q=np.zeros((N,), dtype=f32),
        terminated=np.zeros((N,), dtype=np.uint8),
        emb=[np.zeros((N, rb), dtype=np.uint8) for rb in emb_row_bytes],
        r=np.zeros((N,), dtype=f32) if weighted else None,
    )


def reset(tree: Tree) -> None:
    """tree.py:272-278 (in place)."""
    tree.next_free_idx = 0
    tree.parents[:] = NULL_INDEX
    tree.edge_map[:] = NULL_INDEX
    tree.n[:] = 0
    tree.p[:] = 0
    tree.q[:] = 0
    tree.terminated[:] = 0
    for e in tree.emb:
        e[:] = 0
    if tree.r is not None:
        tree.r[:] = 0


def get_child_data(tree: Tree, field_arr: np.ndarray, index: int) -> np.ndarray:
    """tree.py:78-98: gather through edge_map[index]; missing children read NULL_VALUE = 0.
    (The gather at -1 wraps to the last row in JAX; the where() discards it.)"""
    mapping = tree.edge_map[index]
    child = field_arr[mapping]  # numpy also wraps -1
    return np.where(mapping == NULL_INDEX, field_arr.dtype.type(0), child)


def set_root(tree: Tree, root_policy, root_value, root_emb) -> None:
    """mcts.py:363-384 update_root_node (weighted_mcts.py:66-87) + tree.py:135-150 set_root."""
    visited = tree.n[ROOT_INDEX] > 0
    tree.p[ROOT_INDEX] = np.asarray(root_policy, dtype=f32)
    if not visited:
        tree.q[ROOT_INDEX] = f32(root_value)
        tree.n[ROOT_INDEX] = 1
        if tree.r is not None:
            tree.r[ROOT_INDEX] = f32(root_value)
    for k, e in enumerate(tree.emb):
        e[ROOT_INDEX] = root_emb[k]
    tree.next_free_idx = max(int(tree.next_free_idx), 1)  # tree.py:147


# ----------------------------------------------------------------------------------------------------
# action selection (core/evaluators/mcts/action_selection.py)
# ----------------------------------------------------------------------------------------------------
@dataclass
class SearchCfg:
    selector: int = SEL_PUCT
    c: float = 1.0  # action_selection.py:67
    c1: float = 1.25  # action_selection.py:123
    c2: float = 19652.0  # action_selection.py:124
    epsilon: float = 1e-8  # action_selection.py:41
    discount: float = -1.0  # mcts.py:24
    weighted: bool = False
    q_temperature: float = 1.0  # weighted_mcts.py:25
    fma_backup: bool = False
    q_transform: int = 0  # include/tz_abi.h TZ_QT_*: 0 normalize_q_values (action_selection.py:70), 1 identity


def normalize_q_values(q_values, child_n, parent_q, epsilon) -> np.ndarray:
    """action_selection.py:10-32."""
    mn = np.minimum(f32(parent_q), q_values.min())
    mx = np.maximum(f32(parent_q), q_values.max())
    completed = np.where(child_n > 0, q_values, mn).astype(f32)
    denom = np.maximum((mx - mn).astype(f32), f32(epsilon))
    return ((completed - mn).astype(f32) / denom).astype(f32)


def q_transform(kind: int, q_values, child_n, parent_q, epsilon) -> np.ndarray:
    """The registry of q_transform functors (include/tz_abi.h TZ_QT_*), the `q_transform` constructor argument of
    action_selection.py:70,128."""
    if kind == 0:
        return normalize_q_values(q_values, child_n, parent_q, epsilon)
    if kind == 1:  # lambda q, n, parent_q, eps: q
        return q_values
    raise ValueError(f"unknown q_transform {kind}")


def select_action(tree: Tree, index: int, cfg: SearchCfg) -> int:
    """PUCTSelector.__call__ action_selection.py:91-116 / MuZeroPUCTSelector.__call__ :150-177."""
    node_q, node_n, node_p = tree.q[index], tree.n[index], tree.p[index]
    q_values = get_child_data(tree, tree.q, index).astype(f32)
    dq = (q_values * f32(cfg.discount)).astype(f32)
    n_values = get_child_data(tree, tree.n, index).astype(i32)
    qn = q_transform(cfg.q_transform, dq, n_values, node_q, cfg.epsilon)  # self.q_transform(...), :109
    sq = np.sqrt(f32(node_n))
    denom = (n_values + i32(1)).astype(f32)
    if cfg.selector == SEL_PUCT:
        u = (((f32(cfg.c) * node_p).astype(f32) * sq).astype(f32) / denom).astype(f32)  # :112
    else:
        base = ((node_p * sq).astype(f32) / denom).astype(f32)  # :171
        t = ((f32(node_n) + f32(cfg.c2)).astype(f32) + f32(1.0)).astype(f32)
        log_term = (tz_logf((t / f32(cfg.c2)).astype(f32)) + f32(cfg.c1)).astype(f32)  # :172
        u = (base * log_term).astype(f32)
    return int(np.argmax((qn + u).astype(f32)))  # first max, :116


def traverse(tree: Tree, cfg: SearchCfg) -> Tuple[int, int, int]:
    """mcts.py:192-228.  Returns (parent, action, levels) where levels counts selector calls."""
    parent = ROOT_INDEX
    action = select_action(tree, ROOT_INDEX, cfg)
    levels = 1
    while True:
        child = int(tree.edge_map[parent, action])
        if child == NULL_INDEX or tree.terminated[child]:  # cond_fn mcts.py:208-213
            break
        parent = child
        action = select_action(tree, parent, cfg)
        levels += 1
    return parent, action, levels


# ----------------------------------------------------------------------------------------------------
# expansion and backpropagation (mcts.py:174-189, 231-262, 299-360; weighted_mcts.py:90-152)
# ----------------------------------------------------------------------------------------------------
def _backup_q(q, n, value, fma: bool) -> np.float32:
    """mcts.py:322  ((q * n) + value) / (n + 1)."""
    if fma:
        num = f32(np.float64(q) * np.float64(f32(n)) + np.float64(value))  # exact product: one rounding
        # NB double rounding of the fused result is possible in principle; only used for the FMA what-if.
    else:
        num = f32(f32(q * f32(n)) + f32(value))
    return f32(num / f32(n + 1))


def expand(tree: Tree, parent: int, action: int, policy, value, terminated, new_emb, cfg: SearchCfg) -> None:
    """mcts.py:174-187: visit an existing (terminal) child or allocate a new one (tree.py:101-132)."""
    node_idx = int(tree.edge_map[parent, action])
    if node_idx != NULL_INDEX:  # visit_node with overwrite, mcts.py:179 + tree.py:153-166
        tree.q[node_idx] = _backup_q(tree.q[node_idx], tree.n[node_idx], f32(value), cfg.fma_backup)
        tree.n[node_idx] += 1
        tree.p[node_idx] = np.asarray(policy, dtype=f32)
        tree.terminated[node_idx] = np.uint8(bool(terminated))
        for k, e in enumerate(tree.emb):
            e[node_idx] = new_emb[k]
        return
    nfi = int(tree.next_free_idx)
    in_bounds = nfi < tree.capacity  # tree.py:116
    if in_bounds:  # out-of-bounds scatter is dropped by JAX
        tree.parents[nfi] = parent
        tree.n[nfi] = 1
        tree.p[nfi] = np.asarray(policy, dtype=f32)
        tree.q[nfi] = f32(value)
        tree.terminated[nfi] = np.uint8(bool(terminated))
        if tree.r is not None:
            tree.r[nfi] = f32(value)  # weighted_mcts.py:60
        for k, e in enumerate(tree.emb):
            e[nfi] = new_emb[k]
    tree.edge_map[parent, action] = nfi if in_bounds else NULL_INDEX  # tree.py:123,128
    tree.next_free_idx = nfi + 1 if in_bounds else nfi


def backpropagate(tree: Tree, parent: int, value, cfg: SearchCfg) -> None:
    """mcts.py:231-262."""
    node = parent
    v = f32(value)
    d = f32(cfg.discount)
    while node != NULL_INDEX:
        v = f32(v * d)  # mcts.py:247
        tree.q[node] = _backup_q(tree.q[node], tree.n[node], v, cfg.fma_backup)
        tree.n[node] += 1
        node = int(tree.parents[node])


def weighted_backpropagate(tree: Tree, parent: int, cfg: SearchCfg, noise=None) -> None:
    """weighted_mcts.py:90-152.  `noise` = uniform(0, tiebreak_noise) of shape (F,), q_temperature == 0 only."""
    node = parent
    d = f32(cfg.discount)
    while node != NULL_INDEX:
        cq = (get_child_data(tree, tree.q, node).astype(f32) * d).astype(f32)
        cn = get_child_data(tree, tree.n, node).astype(i32)
        nq = normalize_q_values(cq, cn, tree.q[node], FLT_EPS)  # :111
        if cfg.q_temperature > 0:
            vals = tz_powf(nq, f32(1.0 / cfg.q_temperature))  # :115
            logits = np.where(cn > 0, nq, -FLT_MAX).astype(f32)  # :117-119
        else:
            noisy = (nq + np.asarray(noise, dtype=f32)).astype(f32)  # :123-124
            logits = np.full_like(nq, -FLT_MAX)
            logits[int(np.argmax(noisy))] = f32(1.0)  # :128-130
            vals = nq
        w = softmax(logits)  # :135
        qw = canon_sum((w * vals).astype(f32))  # :137
        tree.q[node] = _backup_q(qw, tree.n[node], tree.r[node], cfg.fma_backup)  # :139-142
        tree.n[node] += 1
        node = int(tree.parents[node])


# ----------------------------------------------------------------------------------------------------
# root action (mcts.py:265-296) and value (mcts.py:111-120)
# ----------------------------------------------------------------------------------------------------
def root_action(tree: Tree, temperature: float, noise=None, uniform01=None):
    """Returns (action, policy_weights, visits).  temperature == 0 needs `noise` (F,) = uniform(0,
    tiebreak_noise); temperature > 0 needs the scalar `uniform01` consumed by jax.random.choice:
    r = cumsum(p)[-1] * (1 - u); action = searchsorted(cumsum(p), r)."""
    F = tree.branching_factor
    visits = get_child_data(tree, tree.n, ROOT_INDEX).astype(i32)
    total = int(visits.sum())
    if total > 0:
        pw = (visits.astype(f32) / f32(max(total, 1))).astype(f32)
    else:
        pw = np.full((F,), f32(1.0 / F), dtype=f32)
    if temperature == 0:
        return int(np.argmax((pw + np.asarray(noise, dtype=f32)).astype(f32))), pw, visits
    pwt = tz_powf(pw, f32(1.0 / temperature))
    pwt = (pwt / canon_sum(pwt)).astype(f32)
    cum = np.zeros((F,), dtype=f32)
    acc = f32(0.0)
    for a in range(F):
        acc = f32(acc + pwt[a])
        cum[a] = acc
    rr = f32(cum[-1] * f32(f32(1.0) - f32(uniform01)))
    return int(np.sum(cum < rr)), pw, visits


# ----------------------------------------------------------------------------------------------------
# re-rooting (tree.py:169-269), written with the same gathers / scatters as the reference
# ----------------------------------------------------------------------------------------------------
def _get_translation(tree: Tree, child_index: int):
    """tree.py:169-217."""
    N = tree.capacity
    subtrees = np.arange(N, dtype=i32)
    for _ in range(N - 1):  # fori_loop(0, capacity-1) :198; stopping at the fixed point gives equal labels
        parents_subtrees = np.where(tree.parents != NULL_INDEX, subtrees[tree.parents], 0)
        new = np.where(parents_subtrees > 0, parents_subtrees, subtrees).astype(i32)
        if np.array_equal(new, subtrees):
            break
        subtrees = new
    subtree_idx = tree.edge_map[ROOT_INDEX, child_index]
    retain = subtrees == subtree_idx
    slots = np.arange(N, dtype=i32)
    old_idx = retain * slots
    cumsum = np.cumsum(retain).astype(i32)
    new_next = int(cumsum[-1])
    translation = np.where(retain, retain * (cumsum - 1), NULL_INDEX).astype(i32)
    erase = slots >= new_next
    return old_idx, translation, erase


def get_subtree(tree: Tree, action: int) -> None:
    """tree.py:220-269 (in place on `tree`; the reference returns a new pytree)."""
    old_idx, translation, erase = _get_translation(tree, action)
    new_next = int(translation.max()) + 1  # :232

    def translate(x, null_value=0):
        out = x.copy()
        out[translation] = x[old_idx]  # -1 wraps to the last row, which is always erased below
        out[erase] = null_value
        return out

    def translate_idx(x):
        mapped = np.where(x == NULL_INDEX, NULL_INDEX, translation[x]).astype(i32)
        out = x.copy()
        out[translation] = mapped
        out[erase] = NULL_INDEX
        return out

    tree.parents = translate_idx(tree.parents)
    tree.edge_map = translate_idx(tree.edge_map)
    tree.n = translate(tree.n)
    tree.p = translate(tree.p)
    tree.q = translate(tree.q)
    tree.terminated = translate(tree.terminated)
    tree.emb = [translate(e) for e in tree.emb]
    if tree.r is not None:
        tree.r = translate(tree.r)
    tree.next_free_idx = new_next


def step(tree: Tree, action: int, persist_tree: bool = True) -> None:
    """mcts.py:399-414."""
    if persist_tree:
        get_subtree(tree, action)
    else:
        reset(tree)
