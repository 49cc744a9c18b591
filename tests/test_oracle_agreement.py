"""The two independently written oracles agree bit-for-bit: the literal NumPy restatement of the reference
(oracle/mcts_numpy.py, same gathers / scatters / N-1 label rounds as the reference) and the C restatement
(oracle/tz_oracle.c, forward-sweep re-root, scalar loops), driven step-by-step and tree-major."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import Schedule, assert_trees_equal, check_invariants, run_c_stepwise, run_c_treemajor, run_numpy
from oracle import c_oracle as CO
from oracle import mcts_numpy as M
from oracle import synth_numpy as SN


def G(**kw):
    return SN.SynthGame(**kw)


CASES = {
    "ttt_T0": dict(game=G(F=9, payload_bytes=7, rho256=154, tau1024=40, max_depth=9, seed=1), B=4, N=12, S=30, moves=6, temperature=0.0),
    "ttt_cfg1_shape": dict(game=SN.make_game("tic_tac_toe", 1001), B=3, N=128, S=64, moves=3, temperature=1.0),
    "c4": dict(game=G(F=7, payload_bytes=32, rho256=230, tau1024=12, max_depth=42, seed=2), B=4, N=64, S=40, moves=4, temperature=1.0),
    "othello_weighted": dict(game=G(F=65, payload_bytes=0, rho256=38, tau1024=6, max_depth=60, seed=3), B=2, N=50, S=30, moves=3,
                             temperature=1.0, weighted=True),
    "weighted_T05": dict(game=G(F=33, payload_bytes=5, rho256=60, tau1024=6, max_depth=60, seed=3), B=2, N=50, S=30, moves=3,
                         temperature=0.5, weighted=True, q_temperature=0.5),
    "weighted_qT0": dict(game=G(F=33, payload_bytes=5, rho256=100, tau1024=6, max_depth=60, seed=4), B=2, N=50, S=30, moves=3,
                         temperature=1.0, weighted=True, q_temperature=0.0),
    "go_muzero": dict(game=G(F=82, payload_bytes=48, rho256=205, tau1024=2, max_depth=120, seed=5), B=2, N=80, S=60, moves=2,
                      temperature=1.0, selector=1),
    "g2048_pos_discount": dict(game=G(F=4, payload_bytes=16, rho256=218, tau1024=4, max_depth=200, seed=6), B=4, N=40, S=30,
                               moves=5, temperature=1.0, discount=1.0, dirichlet=False),
    "no_persist": dict(game=G(F=7, payload_bytes=3, rho256=230, tau1024=100, max_depth=6, seed=7), B=3, N=64, S=40, moves=6,
                       temperature=1.0, persist_tree=False),
    "deep_discount09": dict(game=G(F=5, payload_bytes=3, rho256=230, tau1024=0, max_depth=100, seed=8), B=2, N=200, S=150,
                            moves=2, temperature=1.0, discount=0.9, c=2.5),
    "fma": dict(game=G(F=7, payload_bytes=16, rho256=230, tau1024=12, max_depth=42, seed=9), B=4, N=64, S=48, moves=3,
                temperature=1.0, fma_backup=True),
    "full_tree": dict(game=G(F=6, payload_bytes=4, rho256=200, tau1024=10, max_depth=30, seed=11), B=4, N=8, S=40, moves=4,
                      temperature=1.0),
}


@pytest.mark.parametrize("name", list(CASES))
def test_numpy_vs_c_stepwise(name):
    s = Schedule(**CASES[name])
    a, b = run_numpy(s, snapshots=True), run_c_stepwise(s, snapshots=True)
    assert np.array_equal(a.actions, b.actions) and np.array_equal(a.pw, b.pw)
    for m, (x, y) in enumerate(zip(a.snapshots, b.snapshots)):
        assert_trees_equal(x, y, f"{name}: after the search of move {m}")
    assert_trees_equal(a.arrays, b.arrays, name)
    check_invariants(b.arrays)


@pytest.mark.parametrize("name", [n for n in CASES if n != "weighted_qT0"])
def test_c_treemajor_vs_stepwise(name):
    s = Schedule(**CASES[name])
    a, b = run_c_stepwise(s), run_c_treemajor(s, nthreads=3)
    assert np.array_equal(a.actions, b.actions) and np.array_equal(a.pw, b.pw)
    assert_trees_equal(a.arrays, b.arrays, name)
    assert np.array_equal(a.stats, b.stats)


def test_env_offset_makes_shards_independent_of_batch_position():
    """Sharding contract (core/common.py:12-29): tree b of shard r equals tree r*B/D + b of the unsharded run."""
    g = G(F=7, payload_bytes=8, rho256=230, tau1024=12, max_depth=42, seed=21)
    B, D = 8, 2
    full = Schedule(game=g, B=B, N=48, S=24, moves=3)
    ref = run_c_treemajor(full)
    for r in range(D):
        sl = slice(r * B // D, (r + 1) * B // D)
        shard = Schedule(game=g, B=B // D, N=48, S=24, moves=3, env_offset=r * B // D)
        shard.dir_noise, shard.root_noise, shard.uniform01 = (np.ascontiguousarray(x[:, sl]) for x in
                                                              (full.dir_noise, full.root_noise, full.uniform01))
        got = run_c_treemajor(shard)
        assert np.array_equal(got.actions, ref.actions[:, sl])
        assert_trees_equal(got.arrays, {k: v[sl] for k, v in ref.arrays.items()}, f"shard {r}")


@settings(max_examples=25, deadline=None)
@given(F=st.integers(1, 40), N=st.integers(2, 40), S=st.integers(1, 40), seed=st.integers(0, 10 ** 6),
       rho=st.integers(0, 256), tau=st.integers(0, 300), payload=st.integers(0, 9), weighted=st.booleans(),
       discount=st.sampled_from([-1.0, 1.0, 0.5, -0.9]), temperature=st.sampled_from([0.0, 1.0, 0.25]))
def test_random_schedules_numpy_vs_c(F, N, S, seed, rho, tau, payload, weighted, discount, temperature):
    s = Schedule(game=G(F=F, payload_bytes=payload, rho256=rho, tau1024=tau, max_depth=12, seed=seed), B=2, N=N, S=S, moves=3,
                 temperature=temperature, weighted=weighted, discount=discount, seed=seed)
    a, b = run_numpy(s), run_c_stepwise(s)
    assert np.array_equal(a.actions, b.actions) and np.array_equal(a.pw, b.pw)
    assert_trees_equal(a.arrays, b.arrays, "random schedule")
    check_invariants(b.arrays)


def test_math_numpy_equals_c_and_is_accurate():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-90, 89, 20000), [-87.0, -86.0, 0.0, -0.0, 88.0, 1e-30, -1e-30]]).astype(np.float32)
    e_c, e_np = CO.math("exp", x), M.tz_expf(x)
    assert np.array_equal(e_c, e_np)
    ok = (x > -86) & (x < 88)
    rel = np.abs(e_c[ok].astype(np.float64) / np.exp(x[ok].astype(np.float64)) - 1)
    assert rel.max() < 4e-7
    y = np.concatenate([np.exp(rng.uniform(-80, 80, 20000)), [1.0, 0.5, 2.0, np.finfo(np.float32).tiny, 3.4e38]]).astype(np.float32)
    l_c, l_np = CO.math("log", y), M.tz_logf(y)
    assert np.array_equal(l_c, l_np)
    assert np.abs(l_c.astype(np.float64) - np.log(y.astype(np.float64))).max() < 1e-5
    z = rng.uniform(0, 1, 5000).astype(np.float32)
    for t in (1.0, 2.0, 4.0, 0.5, 1 / 3):
        p_c, p_np = CO.math("pow", z, float(np.float32(t))), M.tz_powf(z, np.float32(t))
        assert np.array_equal(p_c, p_np)
        np.testing.assert_allclose(p_c, z.astype(np.float64) ** float(np.float32(t)), rtol=2e-5, atol=1e-30)
    assert np.array_equal(M.tz_powf(z, 1.0), z)  # temperature 1 is the identity, exactly


def test_forward_sweep_equals_label_propagation_fixed_point():
    """C re-root (one forward sweep, uses parents[i] < i) vs the reference's N-1 propagation rounds, on random trees."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        N, F = int(rng.integers(2, 60)), int(rng.integers(1, 6))
        t = M.init_tree(N, F, [3])
        k = int(rng.integers(1, N + 1))
        t.next_free_idx = k
        t.n[:k] = rng.integers(1, 50, k)
        t.q[:k] = rng.standard_normal(k).astype(np.float32)
        t.p[:k] = rng.random((k, F), dtype=np.float32)
        t.terminated[:k] = rng.integers(0, 2, k)
        t.emb[0][:k] = rng.integers(0, 256, (k, 3))
        for i in range(1, k):
            free = [(p, a) for p in range(i) for a in range(F) if t.edge_map[p, a] < 0]
            if not free:
                t.next_free_idx = i
                for arr in (t.n, t.q, t.p, t.terminated, t.emb[0]):
                    arr[i:] = 0
                break
            p, a = free[int(rng.integers(len(free)))]
            t.parents[i], t.edge_map[p, a] = p, i
        from test_oracle_known_answers import assert_host_equals, to_host
        for action in range(F):
            h = to_host([t])
            u = t.copy()
            M.get_subtree(u, action)
            CO.reroot(h, np.array([action], np.int32))
            assert_host_equals(h, 0, u)
