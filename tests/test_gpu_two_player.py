"""GPU parity of the two-evaluator game loop (turbozero_b200.two_player_game_step, core/common.py:146-232) against
fixtures produced by the reference's own source (tests/golden/make_golden_two_player.py)."""
import os

import numpy as np
import pytest

from oracle import synth_numpy as SN

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "two_player_c4": (dict(F=7, payload_bytes=8, rho256=230, tau1024=90, max_depth=9, seed=41), dict(N=40, S=24), dict(N=24, S=10, c=1.5), 10),
    "two_player_ttt_T0": (dict(F=9, payload_bytes=0, rho256=154, tau1024=60, max_depth=7, seed=42), dict(N=16, S=20, temperature=0.0),
                          dict(N=32, S=12), 8),
}


def _play_group(name, games, p1_first, fx):
    import torch
    import turbozero_b200 as tz
    from standin.synthetic import SyntheticEnv, SyntheticGame, make_synthetic_evaluator

    gkw, e1, e2, max_steps = CASES[name]
    g = SN.SynthGame(**gkw)
    game = SyntheticGame(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    G = len(games)

    def make_ev(kw):
        return make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(c=kw.get("c", 1.0)), max_nodes=kw["N"],
                                        num_iterations=kw["S"], temperature=kw.get("temperature", 1.0))

    ev1, ev2 = make_ev(e1), make_ev(e2)
    # envs with the fixture's game ids: env b of the fixture starts from init_h(b, 0)
    env = SyntheticEnv(game, G)
    core = np.zeros((G, 4), np.int32)
    for i, b in enumerate(games):
        core[i, 0] = np.uint32(g.init_h(int(b), 0)).view(np.int32) if hasattr(np.uint32(0), "view") else g.init_h(int(b), 0)
    env.state["core"].copy_(torch.from_numpy(core))
    if g.payload_bytes > 0:
        pay = np.stack([g.make_emb(g.init_h(int(b), 0), 0, 0)[1] for b in games])
        env.state["payload"].copy_(torch.from_numpy(pay))

    def metadata_of(core_t, terminated):
        c = core_t.cpu().numpy()
        rew = np.array([[g.reward(int(h) & 0xFFFFFFFF)] for h in c[:, 0]], np.float32)
        return tz.StepMetadata(rewards=torch.from_numpy(np.concatenate([rew, rew], 1)).cuda(),
                               action_mask=torch.ones((G, g.F), dtype=torch.bool, device="cuda"),
                               terminated=terminated, cur_player_id=core_t[:, 2].clone(), step=core_t[:, 1].clone())

    def env_step_fn(state, action):
        # the stand-in re-initialises a finished env in place; the rewards of the step are those of the state it reached
        c = state["core"].cpu().numpy()
        a = action.cpu().numpy()
        h2 = [g.step_h(int(h) & 0xFFFFFFFF, int(x)) for h, x in zip(c[:, 0], a)]
        rew = np.array([[g.reward(h)] for h in h2], np.float32)
        state, md = env.env_step_fn(state, action)
        return state, md.replace(rewards=torch.from_numpy(np.concatenate([rew, rew], 1)).cuda(),
                                 cur_player_id=state["core"][:, 2].clone(), step=torch.from_numpy(c[:, 1] + 1).cuda().to(torch.int32),
                                 terminated=md.terminated.bool())

    tmpl = game.template_embedding()
    dev = "cuda"
    st = tz.TwoPlayerGameState(
        key=None, env_state=env.state, env_state_metadata=metadata_of(env.state["core"], torch.zeros((G,), dtype=torch.bool, device=dev)),
        p1_eval_state=ev1.init_batched(G, tmpl), p2_eval_state=ev2.init_batched(G, tmpl),
        p1_value_estimate=torch.zeros((G,), device=dev), p2_value_estimate=torch.zeros((G,), device=dev),
        outcomes=torch.zeros((G, 2), device=dev), completed=torch.zeros((G,), dtype=torch.bool, device=dev))
    T = fx["ref_actions"].shape[0]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for t in range(T):
        use_p1 = p1_first == (t % 2 == 0)
        before = st.completed.cpu().numpy()
        st, action = tz.two_player_game_step(st, ev1, ev2, None, env_step_fn, None, use_p1, max_steps, return_action=True,
                                             leaf_fn=game.leaf_fn, dirichlet_noise=d(fx["in_dir_noise"][t, games]),
                                             root_noise=d(fx["in_root_noise"][t, games]), uniform01=d(fx["in_uniform01"][t, games]))
        live = ~before
        assert np.array_equal(before, np.concatenate([[False] * 0, fx["ref_completed"][t - 1, games]]) if t else np.zeros(G, bool))
        assert np.array_equal(action.cpu().numpy()[live], fx["ref_actions"][t, games][live]), (name, t)
        assert np.array_equal(st.p1_value_estimate.cpu().numpy(), fx["ref_p1_value"][t, games]), (name, t)
        assert np.array_equal(st.p2_value_estimate.cpu().numpy(), fx["ref_p2_value"][t, games]), (name, t)
        assert np.array_equal(st.completed.cpu().numpy(), fx["ref_completed"][t, games]), (name, t)
        assert np.array_equal(st.p1_eval_state.next_free_idx.cpu().numpy()[live], fx["ref_p1_nfi"][t, games][live]), (name, t)
        assert np.array_equal(st.p2_eval_state.next_free_idx.cpu().numpy()[live], fx["ref_p2_nfi"][t, games][live]), (name, t)
    assert np.array_equal(st.outcomes.cpu().numpy(), fx["ref_outcomes"][games])


@pytest.mark.parametrize("name", list(CASES))
def test_two_player_game_step_matches_reference_fixture(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    first = fx["in_p1_first"]
    for p1_first in (True, False):
        games = np.nonzero(first == p1_first)[0]
        assert len(games) > 0
        _play_group(name, games, p1_first, fx)


def test_two_player_game_driver_runs_to_completion():
    """two_player_game (common.py:235-367) end to end with random noise: games complete, outcomes are frozen once set."""
    import torch
    import turbozero_b200 as tz
    from standin.synthetic import SyntheticEnv, SyntheticGame, make_synthetic_evaluator

    game = SyntheticGame(7, 8, 230, 200, 6, 77)
    G = 32
    mk = lambda S: make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=32, num_iterations=S)
    env = SyntheticEnv(game, G)

    def env_init_fn(key, n):
        return env.state, env.metadata().replace(cur_player_id=env.state["core"][:, 2].clone())

    def env_step_fn(state, action):
        state, md = env.env_step_fn(state, action)
        r = (state["core"][:, 0] & 1).to(torch.float32) * 2 - 1
        return state, md.replace(rewards=torch.stack([r, -r], 1), cur_player_id=state["core"][:, 2].clone(), terminated=md.terminated.bool())

    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    outcomes, state, frames, p_ids = tz.two_player_game(gen, mk(16), mk(8), None, None, env_step_fn, env_init_fn, 12, num_games=G,
                                                        p1_first=True, frames=True, leaf_fn=game.leaf_fn)
    assert bool(state.completed.all()) and outcomes.shape == (G, 2) and len(frames) == 13
    done = torch.stack([f.completed for f in frames])
    assert bool((done[1:] >= done[:-1]).all())  # monotone
    for a, b in zip(frames[:-1], frames[1:]):   # outcomes never change once a game has completed
        assert bool((a.outcomes[a.completed] == b.outcomes[a.completed]).all())
    assert bool((outcomes.abs() == 1).all()) and bool((outcomes.sum(1) == 0).all())
