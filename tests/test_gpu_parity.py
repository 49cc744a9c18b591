"""GPU parity: the CUDA path (through the C-ABI) against the C oracle on identical seeded inputs.
Integers / bytes bit-exact, floats equal (== ; the only slack is the sign of zero)."""
import numpy as np
import pytest

from helpers import (Schedule, assert_trees_equal, check_invariants, run_c_stepwise, run_c_treemajor, run_cuda_api,
                     run_cuda_selfplay)
from oracle import synth_numpy as SN

pytestmark = pytest.mark.gpu


def G(**kw):
    return SN.SynthGame(**kw)


CASES = {
    "ttt_T0": dict(game=G(F=9, payload_bytes=7, rho256=154, tau1024=40, max_depth=9, seed=1), B=5, N=12, S=30, moves=6, temperature=0.0),
    "ttt_cfg1": dict(game=SN.make_game("tic_tac_toe", 1001), B=32, N=128, S=64, moves=5, temperature=1.0),
    "c4": dict(game=G(F=7, payload_bytes=32, rho256=230, tau1024=12, max_depth=42, seed=2), B=9, N=64, S=40, moves=4, temperature=1.0),
    "othello_weighted": dict(game=G(F=65, payload_bytes=0, rho256=38, tau1024=6, max_depth=60, seed=3), B=4, N=50, S=30, moves=3,
                             temperature=1.0, weighted=True),
    "othello_weighted_T05": dict(game=G(F=65, payload_bytes=5, rho256=38, tau1024=6, max_depth=60, seed=3), B=4, N=50, S=30,
                                 moves=3, temperature=0.5, weighted=True, q_temperature=0.5),
    "weighted_qT0": dict(game=G(F=33, payload_bytes=5, rho256=100, tau1024=6, max_depth=60, seed=4), B=3, N=50, S=30, moves=3,
                         temperature=1.0, weighted=True, q_temperature=0.0),
    "go_muzero": dict(game=G(F=82, payload_bytes=48, rho256=205, tau1024=2, max_depth=120, seed=5), B=3, N=80, S=60, moves=3,
                      temperature=1.0, selector=1),
    "g2048_pos_discount": dict(game=G(F=4, payload_bytes=16, rho256=218, tau1024=4, max_depth=200, seed=6), B=6, N=40, S=30,
                               moves=5, temperature=1.0, discount=1.0, dirichlet=False),
    "no_persist": dict(game=G(F=7, payload_bytes=3, rho256=230, tau1024=100, max_depth=6, seed=7), B=4, N=64, S=40, moves=6,
                       temperature=1.0, persist_tree=False),
    "deep_discount09": dict(game=G(F=5, payload_bytes=3, rho256=230, tau1024=0, max_depth=100, seed=8), B=3, N=200, S=150,
                            moves=3, temperature=1.0, discount=0.9, c=2.5),
    "fma": dict(game=G(F=7, payload_bytes=16, rho256=230, tau1024=12, max_depth=42, seed=9), B=8, N=64, S=48, moves=4,
                temperature=1.0, fma_backup=True),
    "wide_F300": dict(game=G(F=300, payload_bytes=20, rho256=128, tau1024=2, max_depth=50, seed=10), B=3, N=70, S=50, moves=3,
                      temperature=1.0),
    "very_deep": dict(game=G(F=2, payload_bytes=4, rho256=0, tau1024=0, max_depth=1000, seed=12), B=3, N=120, S=100, moves=2,
                      temperature=1.0, discount=0.97),
}


def _compare(a, b, what):
    assert np.array_equal(a.actions, b.actions), what
    assert np.array_equal(a.pw, b.pw), what
    assert_trees_equal(a.arrays, b.arrays, what)


@pytest.mark.parametrize("name", list(CASES))
def test_api_fused_vs_oracle(name):
    s = Schedule(**CASES[name])
    ref = run_c_stepwise(s, snapshots=True)
    got = run_cuda_api(s, fused=True, snapshots=True)
    for m, (x, y) in enumerate(zip(ref.snapshots, got.snapshots)):
        assert_trees_equal(x, y, f"{name}: after search of move {m}")
    _compare(ref, got, name)
    assert np.array_equal(ref.stats, got.stats)
    check_invariants(got.arrays)


@pytest.mark.parametrize("name", ["ttt_T0", "c4", "othello_weighted", "weighted_qT0", "very_deep"])
def test_api_unfused_vs_oracle(name):
    s = Schedule(**CASES[name])
    _compare(run_c_stepwise(s), run_cuda_api(s, fused=False), name)


@pytest.mark.parametrize("name", [n for n in CASES if n != "weighted_qT0"])
@pytest.mark.parametrize("use_path", [True, False])
def test_c_search_loop_vs_oracle(name, use_path):
    s = Schedule(**CASES[name])
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, use_path=use_path), name)


@pytest.mark.parametrize("name", ["ttt_cfg1", "c4", "othello_weighted"])
def test_graph_replay_vs_oracle(name):
    s = Schedule(**CASES[name])
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), name)


@pytest.mark.parametrize("name", ["ttt_cfg1", "c4", "othello_weighted", "go_muzero"])
@pytest.mark.parametrize("graph", [False, True])
def test_stream_pipelined_slices_vs_oracle(name, graph):
    """The env batch split into 3 uneven slices searched concurrently on 3 streams gives the same trees."""
    s = Schedule(**CASES[name])
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=graph, pipelines=3), name)


@pytest.mark.parametrize("name", ["ttt_T0", "c4", "othello_weighted", "go_muzero", "very_deep", "no_persist", "wide_F300"])
@pytest.mark.parametrize("graph", [False, True])
def test_programmatic_launch_vs_oracle(name, graph):
    """TzSearchCfg.programmatic: per-simulation launches overlap the leaf kernel before them (tree state is read before
    griddepcontrol.wait, leaf results after it) -- same trees, bit for bit."""
    s = Schedule(**CASES[name], programmatic=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=graph), name)


@pytest.mark.parametrize("name", ["c4", "go_muzero", "g2048_pos_discount", "othello_weighted"])
def test_programmatic_bit0_only_vs_oracle(name):
    """TzSearchCfg.programmatic = 1: the search is launched programmatically but does not signal its dependents early (the mode
    bench.py uses for the go_9x9 and 2048 shapes) -- same trees."""
    s = Schedule(**CASES[name], programmatic=1)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), name)


@pytest.mark.parametrize("name", ["c4", "g2048_pos_discount", "ttt_T0"])
@pytest.mark.parametrize("programmatic", [False, True])
def test_large_batch_several_waves_vs_oracle(name, programmatic):
    """More trees than the SMs hold at once (6000 > 2 * 9 * 148): the per-simulation kernel runs in several waves -- same
    trees as the oracle."""
    c = dict(CASES[name])
    c.update(B=6000, moves=2)
    s = Schedule(**c, programmatic=programmatic)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), f"{name}, 6000 trees")


@pytest.mark.parametrize("name", ["c4", "othello_weighted_T05"])
def test_programmatic_launch_python_loop_vs_oracle(name):
    s = Schedule(**CASES[name], programmatic=True)
    _compare(run_c_stepwise(s), run_cuda_api(s, fused=True), name)


def test_connect_four_full_size_programmatic():
    s = Schedule(game=SN.make_game("connect_four", 2002), B=1024, N=256, S=128, moves=4, temperature=1.0, programmatic=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), "connect_four full, programmatic launches")


DEEP = {
    # paths far deeper than the 32-level ring: narrow (lane-per-level kernels) and wide (82-way, 3 register chunks) trees
    "deep_narrow": dict(game=G(F=2, payload_bytes=4, rho256=0, tau1024=0, max_depth=1000, seed=12), B=5, N=160, S=140, moves=3,
                        temperature=1.0, discount=0.97),
    "deep_wide_go": dict(game=G(F=82, payload_bytes=24, rho256=1, tau1024=0, max_depth=600, seed=31), B=4, N=220, S=180, moves=3,
                         temperature=1.0),
    "deep_muzero_pos": dict(game=G(F=12, payload_bytes=0, rho256=2, tau1024=0, max_depth=600, seed=32), B=4, N=150, S=130, moves=3,
                            temperature=1.0, discount=1.0, selector=1, dirichlet=False),
}


@pytest.mark.parametrize("name", list(DEEP))
@pytest.mark.parametrize("use_spill", [True, False])
@pytest.mark.parametrize("programmatic", [False, True])
def test_deep_paths_vs_oracle(name, use_spill, programmatic):
    """Backups through paths of 60-150 levels: with the spilled path record (deep_windows: 32 levels per round trip, decisions
    recomputed) and without it (parents[] chase) -- same trees as the oracle either way."""
    s = Schedule(**DEEP[name], programmatic=programmatic)
    ref = run_c_treemajor(s)
    assert ref.stats[:, 0].sum() / max(ref.stats[:, 1].sum(), 1) > 15, "the schedule is meant to produce deep paths"
    _compare(ref, run_cuda_selfplay(s, use_spill=use_spill), name)


def test_reroot_one_table_at_a_time_fallback():
    """Rows too wide for the all-tables staging area (large batch -> 16 KB stage, 20 KB embedding rows) take k_reroot."""
    s = Schedule(game=G(F=3, payload_bytes=20000, rho256=230, tau1024=20, max_depth=12, seed=21), B=2100, N=6, S=5, moves=3,
                 temperature=1.0)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s), "wide rows")


def test_reroot_odd_row_sizes():
    """Embedding leaves whose row size is not a multiple of 4 go through the byte path of the all-tables re-root."""
    s = Schedule(game=G(F=6, payload_bytes=13, rho256=200, tau1024=30, max_depth=20, seed=22), B=37, N=40, S=30, moves=5,
                 temperature=1.0)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s), "odd rows")


def test_connect_four_full_size_pipelined():
    s = Schedule(game=SN.make_game("connect_four", 2001), B=1024, N=256, S=128, moves=3, temperature=1.0)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True, pipelines=8), "connect_four full, 8 pipelines")


def test_connect_four_full_size():
    """BASELINE.json configs[1] at full size: 1024 envs x 128 simulations, N = 256, subtree persistence on."""
    s = Schedule(game=SN.make_game("connect_four", 2000), B=1024, N=256, S=128, moves=4, temperature=1.0)
    ref = run_c_treemajor(s)
    got = run_cuda_selfplay(s, graph=True)
    _compare(ref, got, "connect_four full")
    check_invariants({k: v[:64] for k, v in got.arrays.items()})


EDGE = {
    # one tree, one action, odd capacity: every walk is a chain; the tree fills up (tree.py:116-131) and keeps backing up
    "F1_N7_B1": dict(game=G(F=1, payload_bytes=1, rho256=0, tau1024=0, max_depth=1000, seed=51), B=1, N=7, S=12, moves=3, temperature=1.0),
    # capacity 1: only the root ever exists; every expansion is dropped
    "N1": dict(game=G(F=5, payload_bytes=3, rho256=200, tau1024=0, max_depth=50, seed=52), B=3, N=1, S=6, moves=3, temperature=1.0),
    # capacity 2 and an odd batch
    "N2_B7": dict(game=G(F=4, payload_bytes=0, rho256=200, tau1024=30, max_depth=50, seed=53), B=7, N=2, S=9, moves=4, temperature=0.0),
    # two register chunks with one action in the second (F = 33), odd capacity (unaligned best-table rows)
    "F33_N33": dict(game=G(F=33, payload_bytes=9, rho256=120, tau1024=10, max_depth=40, seed=54), B=5, N=33, S=40, moves=3, temperature=1.0),
    # the widest dispatch (16 chunks)
    "F512": dict(game=G(F=512, payload_bytes=4, rho256=64, tau1024=4, max_depth=30, seed=55), B=2, N=24, S=30, moves=2, temperature=1.0),
    # a single simulation per move, and none at all (root only: uniform policy weights, mcts.py:279)
    "S1": dict(game=G(F=6, payload_bytes=2, rho256=200, tau1024=20, max_depth=30, seed=56), B=4, N=16, S=1, moves=5, temperature=1.0),
    "S0": dict(game=G(F=6, payload_bytes=2, rho256=200, tau1024=20, max_depth=30, seed=57), B=4, N=16, S=0, moves=3, temperature=1.0),
    # weighted backup on a two-action tree with odd capacity and argmax weights
    "weighted_F2_qT0": dict(game=G(F=2, payload_bytes=0, rho256=256, tau1024=0, max_depth=60, seed=58), B=3, N=45, S=50, moves=2,
                            temperature=1.0, weighted=True, q_temperature=0.0),
}


@pytest.mark.parametrize("name", list(EDGE))
def test_edge_shapes_python_api_vs_oracle(name):
    s = Schedule(**EDGE[name])
    _compare(run_c_stepwise(s), run_cuda_api(s, fused=True), name)


@pytest.mark.parametrize("name", [n for n in EDGE if n != "weighted_F2_qT0"])
@pytest.mark.parametrize("graph", [False, True])
def test_edge_shapes_c_loop_vs_oracle(name, graph):
    s = Schedule(**EDGE[name], programmatic=graph)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=graph), name)


def test_othello_weighted_full_size():
    """BASELINE.json configs[2] per-GPU share at 8 GPUs: 512 envs x 200 simulations, N = 400, WeightedMCTS backup."""
    s = Schedule(game=SN.make_game("othello", 3000), B=512, N=400, S=200, moves=2, temperature=1.0, weighted=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), "othello weighted full")


def test_2048_full_size_positive_discount():
    """BASELINE.json configs[4] per-GPU share: 2048 envs x 100 simulations, N = 200, discount +1, programmatic launches."""
    s = Schedule(game=SN.make_game("2048", 5000), B=2048, N=200, S=100, moves=3, temperature=1.0, discount=1.0, programmatic=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), "2048 full")


def test_go_9x9_full_depth_reduced_batch():
    """BASELINE.json configs[3] tree shape at full depth (N = 1600, 800 simulations, 82-way, 4 KB embedding rows, paths far
    beyond the 32-level ring) on 48 envs -- the oracle needs seconds for that; the full 1024-env batch is checked through
    its invariants below."""
    s = Schedule(game=SN.make_game("go_9x9", 4000), B=48, N=1600, S=800, moves=2, temperature=1.0)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=False), "go_9x9 full depth")


def test_go_9x9_full_size_invariants():
    """configs[3] per-GPU share (1024 envs x 800 simulations): structural invariants of the reference's trees, the derived
    tables against a from-scratch rebuild, the best-table against a re-evaluation of the selector (run_cuda_selfplay checks
    the last two), and visit-count conservation: n[root] = 1 + simulations landed below it."""
    s = Schedule(game=SN.make_game("go_9x9", 4001), B=1024, N=1600, S=800, moves=1, temperature=1.0, persist_tree=False)
    got = run_cuda_selfplay(s, graph=True)  # persist_tree=False: the tree is reset after the move ...
    assert (got.arrays["next_free_idx"] == 0).all() and (got.arrays["parents"] == -1).all() and not got.arrays["n"].any()
    assert np.allclose(got.pw.sum(-1), 1.0, atol=1e-5) and (got.pw >= 0).all()  # ... and the move's policy weights are visit shares
    s2 = Schedule(game=SN.make_game("go_9x9", 4001), B=1024, N=1600, S=800, moves=1, temperature=1.0)
    got2 = run_cuda_selfplay(s2, graph=True)
    check_invariants({k: v[:96] for k, v in got2.arrays.items()})
    assert np.array_equal(got.actions, got2.actions) and np.array_equal(got.pw, got2.pw)  # persistence does not change move 0


def test_division_sequence_is_ieee_exact():
    """The select kernel's straight-line division (div_core) equals div.rn bit-for-bit on 2^30 operand pairs."""
    import torch
    from turbozero_b200 import _abi

    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2):
        _abi.check(_abi.lib().tz_selftest_div(1 << 29, seed, bad.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest")
    assert int(bad.item()) == 0


# ---- round 2: the shapes bench.py times (BASELINE.json configs[2..4] per-GPU shares), against the oracle ----------------
@pytest.mark.parametrize("envs", [1024, 2048])
def test_othello_weighted_2_and_4_gpu_shares(envs):
    """BASELINE.json configs[2] sharded over 4 / 2 GPUs: 1024 / 2048 envs per GPU x 200 simulations, N = 400, WeightedMCTS."""
    s = Schedule(game=SN.make_game("othello", 3001), B=envs, N=400, S=200, moves=2, temperature=1.0, weighted=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), f"othello weighted, {envs} envs")


def test_go_9x9_full_depth_256_envs():
    """BASELINE.json configs[3] tree shape at full depth (N = 1600, 800 simulations, 82-way, 4 KB embedding rows) on a quarter
    of the per-GPU batch, against the multi-threaded C oracle, through a replayed CUDA graph like the bench."""
    s = Schedule(game=SN.make_game("go_9x9", 4002), B=256, N=1600, S=800, moves=2, temperature=1.0)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), "go_9x9 full depth, 256 envs")


def test_2048_step_with_replay_update_at_bench_shape():
    """The configs[4] step exactly as bench.py runs it on one GPU's share: 2048 envs x 100 simulations (N = 200, discount +1)
    FOLLOWED IN THE SAME STEP by the replay-buffer update (tz_replay_collect fed from the move's outputs), three moves in a
    replayed CUDA graph.  Trees / actions / policy weights against the C oracle; the buffer against oracle/replay_numpy.py
    fed with the same per-move records."""
    import torch
    import turbozero_b200 as tz
    from helpers import make_cuda_evaluator
    from oracle import replay_numpy as RN
    from standin.synthetic import SyntheticGame, SyntheticSelfPlay

    B, F, cap, moves = 2048, 4, 16, 3
    s = Schedule(game=SN.make_game("2048", 5001), B=B, N=200, S=100, moves=moves, temperature=1.0, discount=1.0, programmatic=True)
    ref = run_c_treemajor(s)
    g = s.game
    game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    ev = make_cuda_evaluator(s, game)
    sp = SyntheticSelfPlay(game, ev, B, dirichlet=True)
    rb = tz.EpisodeReplayBuffer(capacity=cap)
    obs0 = sp.state["core"].to(torch.float32)
    rstate = rb.init(B, tz.BaseExperience(reward=torch.zeros((1,)), policy_weights=torch.zeros((F,)),
                                          policy_mask=torch.zeros((F,), dtype=torch.bool), observation_nn=obs0[0].cpu(),
                                          cur_player_id=torch.zeros((), dtype=torch.int32)))
    x_obs, x_mask = torch.empty_like(obs0), torch.ones((B, F), dtype=torch.bool, device="cuda")
    x_rew0, x_rew = torch.zeros((B, 1), device="cuda"), torch.empty((B, 1), device="cuda")
    x_player = torch.zeros((B,), dtype=torch.int32, device="cuda")
    x_trunc = torch.zeros((B,), dtype=torch.uint8, device="cuda")

    def one_move():
        x_obs.copy_(sp.state["core"])
        sp.move()
        x_rew.copy_(sp.reset_flag.unsqueeze(1))
        exp = tz.BaseExperience(reward=x_rew0, policy_weights=sp.policy_weights, policy_mask=x_mask, observation_nn=x_obs,
                                cur_player_id=x_player)
        rb.collect_update(rstate, [exp], x_rew, sp.reset_flag, x_trunc)

    tmpl_np = {"reward": np.zeros((1,), np.float32), "policy_weights": np.zeros((F,), np.float32), "policy_mask": np.zeros((F,), bool),
               "observation_nn": np.zeros((4,), np.float32), "cur_player_id": np.zeros((), np.int32)}
    rs = RN.init(B, cap, tmpl_np)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(cg, stream=side):
            one_move()
    torch.cuda.current_stream().wait_stream(side)
    actions = np.zeros((moves, B), np.int32)
    pw = np.zeros((moves, B, F), np.float32)
    for m in range(moves):
        sp.dir_noise.copy_(torch.from_numpy(s.dir_noise[m]))
        sp.root_noise.copy_(torch.from_numpy(s.root_noise[m]))
        sp.uniform01.copy_(torch.from_numpy(s.uniform01[m]))
        obs_before = sp.state["core"].to(torch.float32).cpu().numpy()
        cg.replay()
        torch.cuda.synchronize()
        actions[m], pw[m] = sp.action.cpu().numpy(), sp.policy_weights.cpu().numpy()
        done = sp.reset_flag.cpu().numpy().astype(bool)
        e = {"reward": np.zeros((B, 1), np.float32), "policy_weights": pw[m], "policy_mask": np.ones((B, F), bool),
             "observation_nn": obs_before, "cur_player_id": np.zeros((B,), np.int32)}
        RN.collect_update(rs, [e], done.astype(np.float32).reshape(B, 1), done, np.zeros((B,), bool), cap)
    assert np.array_equal(ref.actions, actions) and np.array_equal(ref.pw, pw)
    from helpers import tree_to_numpy
    assert_trees_equal(ref.arrays, tree_to_numpy(sp.tree), "2048 step with replay update")
    assert np.array_equal(rstate.populated.cpu().numpy(), rs.populated) and np.array_equal(rstate.has_reward.cpu().numpy(), rs.has_reward)
    assert np.array_equal(rstate.next_idx.cpu().numpy(), rs.next_idx)
    for k in tmpl_np:
        assert np.array_equal(getattr(rstate.buffer, k).cpu().numpy(), rs.buffer[k]), k


def test_masked_reset_touches_only_the_flagged_trees():
    """MCTS.reset(state, mask): Tree.reset (tree.py:272-278) for the flagged trees, the others bit-for-bit untouched."""
    import torch
    from helpers import make_cuda_evaluator, tree_to_numpy
    from standin.synthetic import SyntheticGame

    s = Schedule(**CASES["c4"])
    g = s.game
    game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    ev = make_cuda_evaluator(s, game)
    tree = ev.init_batched(s.B, game.template_embedding())
    state, _ = game.init_states(s.B)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    ev.evaluate(None, tree, state, None, None, None, leaf_fn=game.leaf_fn, root_noise=dev(s.root_noise[0]), uniform01=dev(s.uniform01[0]),
                dirichlet_noise=dev(s.dir_noise[0]))
    before = tree_to_numpy(tree)
    mask = torch.tensor([b % 3 == 1 for b in range(s.B)], device="cuda")
    ev.reset(tree, mask=mask)
    after = tree_to_numpy(tree)
    fresh = tree_to_numpy(ev.init_batched(s.B, game.template_embedding()))
    m = mask.cpu().numpy()
    assert m.any() and not m.all() and (before["next_free_idx"] > 1).all()
    for k in before:
        assert np.array_equal(after[k][m], fresh[k][m]), f"{k}: flagged trees must be empty"
        assert np.array_equal(after[k][~m], before[k][~m]), f"{k}: unflagged trees must be untouched"
    ev.reset(tree)  # and the unmasked form resets everything
    assert_trees_equal(fresh, tree_to_numpy(tree), "full reset")


def test_launch_timeline_records_every_search_and_leaf_launch():
    """TzWork.timeline / tz_synth_set_timeline (what bench.py's roofline leg reads): every launch of a replayed move writes
    its row, rows are ordered in time, and the trees are the oracle's with the record on."""
    import torch
    from standin.synthetic import SyntheticGame, SyntheticSelfPlay
    from helpers import make_cuda_evaluator, tree_to_numpy

    for programmatic in (False, True):
        s = Schedule(**CASES["c4"], programmatic=programmatic)
        g = s.game
        game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
        ev = make_cuda_evaluator(s, game)
        sp = SyntheticSelfPlay(game, ev, s.B, dirichlet=True)
        sp.timeline_begin()
        for m in range(s.moves):
            sp.dir_noise.copy_(torch.from_numpy(s.dir_noise[m]))
            sp.root_noise.copy_(torch.from_numpy(s.root_noise[m]))
            sp.uniform01.copy_(torch.from_numpy(s.uniform01[m]))
            sp.timeline_clear()
            mark = sp.timeline_mark()
            sp.move()
            torch.cuda.synchronize()
            ts, tl = sp.timeline_read(mark, s.S + 1, s.S)
            assert (ts[:, 0] > 0).all() and (ts[:, 2] >= ts[:, 0]).all(), "every search launch wrote first-in / last-out"
            assert (tl[:, 0] > 0).all() and (tl[:, 2] >= tl[:, 1]).all() and (tl[:, 1] >= tl[:, 0]).all()
            assert (ts[1:, 2] > ts[:-1, 2]).all() and (tl[1:, 2] > tl[:-1, 2]).all(), "launches complete in stream order"
            assert (tl[:, 2] > ts[:-1, 0]).all() and (ts[1:, 2] > tl[:, 2]).all(), "leaf s sits between search launches s and s+1"
        sp.timeline_end()
        ref = run_c_treemajor(s)
        assert_trees_equal(ref.arrays, tree_to_numpy(sp.tree), f"timeline on, programmatic={programmatic}")


# ---- a CTA of W warps per tree (TzSearchCfg.sim_warps, k_sim_wide): same trees, bit for bit ---------------------------
WIDE_CASES = ["c4", "othello_weighted", "othello_weighted_T05", "go_muzero", "deep_discount09", "wide_F300", "very_deep", "no_persist",
              "g2048_pos_discount", "fma"]


@pytest.mark.parametrize("name", WIDE_CASES)
@pytest.mark.parametrize("warps", [2, 4, 8])
def test_cta_per_tree_c_loop_vs_oracle(name, warps):
    s = Schedule(**CASES[name], sim_warps=warps)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=(warps == 4)), f"{name}, {warps} warps per tree")


@pytest.mark.parametrize("name", ["c4", "othello_weighted", "weighted_qT0", "go_muzero", "ttt_T0"])
@pytest.mark.parametrize("warps", [2, 4])
def test_cta_per_tree_python_api_vs_oracle(name, warps):
    """Through MCTS.evaluate (one launch per simulation from Python), with the per-move self-tests of run_cuda_api."""
    s = Schedule(**CASES[name], sim_warps=warps)
    ref = run_c_stepwise(s, snapshots=True)
    got = run_cuda_api(s, fused=True, snapshots=True)
    for m, (x, y) in enumerate(zip(ref.snapshots, got.snapshots)):
        assert_trees_equal(x, y, f"{name}: after search of move {m}")
    _compare(ref, got, name)
    assert np.array_equal(ref.stats, got.stats)


@pytest.mark.parametrize("name", ["c4", "othello_weighted", "very_deep"])
def test_cta_per_tree_unfused_and_programmatic(name):
    s = Schedule(**CASES[name], sim_warps=4)
    _compare(run_c_stepwise(s), run_cuda_api(s, fused=False), name + " unfused")
    s = Schedule(**CASES[name], sim_warps=4, programmatic=True)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), name + " programmatic")


@pytest.mark.parametrize("name", list(DEEP))
@pytest.mark.parametrize("warps", [2, 4, 8])
def test_cta_per_tree_deep_paths_vs_oracle(name, warps):
    """Paths of 60-150 levels: several 32 W-level passes (W = 2), the vote over a path longer than one pass."""
    s = Schedule(**DEEP[name], sim_warps=warps)
    _compare(run_c_treemajor(s), run_cuda_selfplay(s), f"{name}, {warps} warps")


def test_cta_per_tree_deep_weighted():
    s = Schedule(game=G(F=40, payload_bytes=8, rho256=2, tau1024=0, max_depth=600, seed=33), B=4, N=160, S=140, moves=3,
                 temperature=1.0, weighted=True, sim_warps=2)
    ref = run_c_treemajor(s)
    assert ref.stats[:, 0].sum() / max(ref.stats[:, 1].sum(), 1) > 15, "the schedule is meant to produce deep paths"
    _compare(ref, run_cuda_selfplay(s), "deep weighted, 2 warps")
    s4 = Schedule(game=s.game, B=4, N=160, S=140, moves=3, temperature=1.0, weighted=True, sim_warps=4)
    _compare(ref, run_cuda_selfplay(s4, graph=True), "deep weighted, 4 warps")


def test_cta_per_tree_rebuilds_a_foreign_path_record():
    """select by the one-warp kernel, expand by the CTA-per-tree kernel (and the other way round): the path record of the
    other form is not trusted; the kernel rebuilds the path from parents[] / edge_map -- same trees as the oracle."""
    import torch
    from helpers import make_cuda_evaluator, tree_to_numpy
    from standin.synthetic import SyntheticGame

    for name in ("c4", "go_muzero", "othello_weighted"):
        s = Schedule(**CASES[name])
        ref = run_c_stepwise(s)
        g = s.game
        game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
        ev = make_cuda_evaluator(s, game)
        tree = ev.init_batched(s.B, game.template_embedding())
        state, episode = game.init_states(s.B, s.env_offset)
        reset_flag = torch.zeros((s.B,), dtype=torch.uint8, device="cuda")
        dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
        actions = np.zeros((s.moves, s.B), np.int32)
        for m in range(s.moves):
            ev.update_root(None, tree, state, None, dirichlet_noise=dev(s.dir_noise[m]))
            for it in range(s.S):
                ev.sim_warps = 1 if (it + m) % 2 == 0 else 4  # traverse with one form ...
                ev.traverse(tree)
                ev.sim_warps = 4 if (it + m) % 2 == 0 else 1  # ... expand + backprop with the other
                sc = tree._scratch
                w, keep = ev._leaf_work(None, tree, sc, None, None, game.leaf_fn, None)
                import ctypes as C
                from turbozero_b200 import _abi
                cfg = ev._cfg()
                _abi.check(_abi.lib().tz_expand_backprop(C.byref(tree.struct()), C.byref(cfg), C.byref(w),
                                                         torch.cuda.current_stream().cuda_stream), "tz_expand_backprop")
            act, _ = ev.sample_root_action(None, tree, root_noise=dev(s.root_noise[m]), uniform01=dev(s.uniform01[m]))
            actions[m] = act.cpu().numpy()
            game.env_step(state, act, episode, reset_flag, s.env_offset)
            ev.step(tree, act, reset_mask=reset_flag)
        assert np.array_equal(ref.actions, actions), name
        assert_trees_equal(ref.arrays, tree_to_numpy(tree), name + ": mixed path-record forms")


def test_cta_per_tree_at_bench_shapes():
    """configs[2] (othello weighted, 512 envs) and configs[3] (go_9x9, 128 envs at full depth) with the library's default
    (4 warps per tree for F > 32) and with 8 / 2 warps."""
    for warps in (0, 8):
        s = Schedule(game=SN.make_game("go_9x9", 4003), B=128, N=1600, S=800, moves=2, temperature=1.0, sim_warps=warps)
        _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), f"go_9x9, sim_warps={warps}")
    for warps in (0, 2):
        s = Schedule(game=SN.make_game("othello", 3002), B=512, N=400, S=200, moves=2, temperature=1.0, weighted=True, sim_warps=warps)
        _compare(run_c_treemajor(s), run_cuda_selfplay(s, graph=True), f"othello weighted, sim_warps={warps}")


@pytest.mark.parametrize("name", ["puct_identity_qtransform", "muzero_puct_arityfix"])
def test_registered_q_transforms_with_programmatic_launches(name):
    """The identity q_transform is compiled for ordinary launches only: asking for programmatic launches with it must still
    give the reference's trees (the library launches that shape ordinarily); normalize_q_values keeps the programmatic form."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as MG

    s, z = MG.load_case(name)
    s.programmatic = True
    for graph in (False, True):
        res = run_cuda_selfplay(s, graph=graph)
        assert np.array_equal(res.actions, z["actions"]) and np.array_equal(res.pw, z["pw"]), name
        assert_trees_equal(MG.split(z, "final"), res.arrays, f"{name}: final trees, programmatic, graph={graph}")
