"""GPU parity of the replay-buffer kernels (include/tz_replay.h, through the Python mirror of
core/memory/replay_memory.py) against the committed reference fixtures and the NumPy oracle.  Bit-exact."""
import numpy as np
import pytest

from oracle import replay_numpy as RN
from test_replay_oracle import CASES, FIELDS, replay_fixture, run_oracle

pytestmark = pytest.mark.gpu


def _to_np(state):
    out = {"next_idx": state.next_idx.cpu().numpy(), "episode_start_idx": state.episode_start_idx.cpu().numpy(),
           "populated": state.populated.cpu().numpy(), "has_reward": state.has_reward.cpu().numpy()}
    for f in FIELDS:
        out["buf_" + f] = getattr(state.buffer, f).cpu().numpy()
    return out


def run_cuda(fx, fused=True):
    import torch
    import turbozero_b200 as tz

    x = {k[3:]: fx[k] for k in fx.files if k.startswith("in_")}
    steps, n_exp, B = x["cur_player_id"].shape
    cap = fx["ref_populated"].shape[1]
    P, F = x["rewards"].shape[2], x["policy_weights"].shape[3]
    buf = tz.EpisodeReplayBuffer(capacity=cap)
    tmpl = tz.BaseExperience(reward=torch.zeros((P,)), policy_weights=torch.zeros((F,)), policy_mask=torch.zeros((F,), dtype=torch.bool),
                             observation_nn=torch.zeros(x["observation_nn"].shape[3:]), cur_player_id=torch.zeros((), dtype=torch.int32))
    st = buf.init(B, tmpl)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for t in range(steps):
        exps = [tz.BaseExperience(observation_nn=dev(x["observation_nn"][t, e]), policy_mask=dev(x["policy_mask"][t, e]),
                                  policy_weights=dev(x["policy_weights"][t, e]), reward=torch.zeros((B, P), device="cuda"),
                                  cur_player_id=dev(x["cur_player_id"][t, e])) for e in range(n_exp)]
        rew, term, trunc = dev(x["rewards"][t]), dev(x["terminated"][t]), dev(x["truncated"][t])
        if fused:
            buf.collect_update(st, exps, rew, term, trunc)
        else:  # the reference's call sequence, one method at a time (train.py:300-340)
            for e in exps:
                buf.add_experience(st, e)
            buf.assign_rewards(st, rew, term)
            buf.truncate(st, trunc)
    return buf, st, x


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("fused", [True, False])
def test_buffer_updates_match_reference_fixture(name, fused):
    fx = replay_fixture(name)
    buf, st, x = run_cuda(fx, fused)
    got = _to_np(st)
    for k, v in got.items():
        assert np.array_equal(v, fx["ref_" + k]), k


@pytest.mark.parametrize("name", CASES)
def test_sample_matches_reference_fixture(name):
    import torch

    fx = replay_fixture(name)
    buf, st, x = run_cuda(fx)
    S = fx["ref_sample_reward"].shape[0]
    smp = buf.sample(st, None, S, gumbel=torch.from_numpy(x["gumbel"]).cuda())
    for f in FIELDS:
        assert np.array_equal(getattr(smp, f).cpu().numpy(), fx["ref_sample_" + f]), f


def test_scores_match_oracle_bitwise():
    import torch

    fx = replay_fixture("replay_single_player")
    buf, st, x = run_cuda(fx)
    s, _ = run_oracle(fx)
    scores = buf.sample_scores(st, torch.from_numpy(x["gumbel"]).cuda()).cpu().numpy()
    w = (s.populated & s.has_reward).reshape(-1)
    p = (w.astype(np.float32) / np.float32(w.sum())).astype(np.float32)
    ref = np.where(w, (-x["gumbel"] - RN._logf(p)).astype(np.float32), np.float32(np.inf))
    assert np.array_equal(scores, ref)


def test_large_random_schedule_vs_oracle():
    """2048-env batch, 3 experiences per step (two transforms), 64-slot rings, odd row sizes."""
    import torch
    import turbozero_b200 as tz

    rng = np.random.default_rng(5)
    B, cap, P, F, steps, n_exp = 2048, 64, 2, 7, 25, 3
    tmpl_np = {"reward": np.zeros((P,), np.float32), "policy_weights": np.zeros((F,), np.float32), "policy_mask": np.zeros((F,), bool),
               "observation_nn": np.zeros((3, 5), np.float32), "cur_player_id": np.zeros((), np.int32)}
    s = RN.init(B, cap, tmpl_np)
    buf = tz.EpisodeReplayBuffer(capacity=cap)
    st = buf.init(B, tz.BaseExperience(**{k: torch.from_numpy(v) for k, v in tmpl_np.items()}))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for t in range(steps):
        exps = [{"observation_nn": rng.standard_normal((B, 3, 5)).astype(np.float32), "policy_mask": rng.random((B, F)) < 0.7,
                 "policy_weights": rng.random((B, F)).astype(np.float32), "reward": np.zeros((B, P), np.float32),
                 "cur_player_id": rng.integers(0, 2, (B,)).astype(np.int32)} for _ in range(n_exp)]
        rew = rng.standard_normal((B, P)).astype(np.float32)
        term, trunc = rng.random(B) < 0.1, rng.random(B) < 0.05
        RN.collect_update(s, exps, rew, term, trunc, cap)
        buf.collect_update(st, [tz.BaseExperience(**{k: dev(v) for k, v in e.items()}) for e in exps], dev(rew), dev(term), dev(trunc))
    got = _to_np(st)
    assert np.array_equal(got["next_idx"], s.next_idx) and np.array_equal(got["episode_start_idx"], s.episode_start_idx)
    assert np.array_equal(got["populated"], s.populated) and np.array_equal(got["has_reward"], s.has_reward)
    for f in FIELDS:
        assert np.array_equal(got["buf_" + f], s.buffer[f]), f
    g = rng.gumbel(size=(B * cap,)).astype(np.float32)
    smp = buf.sample(st, None, 4096, gumbel=dev(g))
    ref = RN.sample(s, g, 4096)
    for f in FIELDS:
        assert np.array_equal(getattr(smp, f).cpu().numpy(), ref[f]), f


def test_collect_step_glue_vs_oracle():
    """turbozero_b200.collect (Trainer.collect, train.py:271-347) over a few moves of the synthetic game with one data
    transform: the experiences it stores describe the PRE-step position, and the buffer equals the oracle's fed with the
    same per-step records."""
    import torch
    import turbozero_b200 as tz
    from standin.synthetic import SyntheticEnv, SyntheticGame, make_synthetic_evaluator

    B, cap, F, moves = 48, 8, 7, 12
    game = SyntheticGame(F, 12, 230, 120, 7, 91)
    ev = make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=24, num_iterations=12)
    env = SyntheticEnv(game, B)
    tree = ev.init_batched(B, game.template_embedding())

    def decorate(md, core):  # non-trivial rewards / players / step counters derived from the env state
        r0 = ((core[:, 0] & 3) - 1).to(torch.float32)
        return md.replace(rewards=torch.stack([r0, -r0], 1), cur_player_id=core[:, 2].clone(), step=core[:, 1].clone(),
                          action_mask=((core[:, 0:1] >> torch.arange(F, device=core.device)) & 1).bool() | (torch.arange(F, device=core.device) == 0))

    def env_step_fn(state, action):
        state, md = env.env_step_fn(state, action)
        return state, decorate(md, state["core"])

    obs_fn = lambda s: s["core"].to(torch.float32) * 0.5
    flip = lambda mask, pw, s: (mask.flip(1), pw.flip(1), s)
    records = []

    class Recorder(tz.EpisodeReplayBuffer):
        def collect_update(self, state, experiences, reward, terminated, truncated):
            records.append(([{f: getattr(e, f).cpu().numpy().copy() for f in FIELDS} for e in experiences],
                            reward.cpu().numpy().copy(), terminated.cpu().numpy().copy(), truncated.cpu().numpy().copy()))
            return super().collect_update(state, experiences, reward, terminated, truncated)

    buf = Recorder(capacity=cap)
    tmpl = tz.BaseExperience(reward=torch.zeros((2,)), policy_weights=torch.zeros((F,)), policy_mask=torch.zeros((F,), dtype=torch.bool),
                             observation_nn=torch.zeros((4,)), cur_player_id=torch.zeros((), dtype=torch.int32))
    state = tz.CollectionState(eval_state=tree, env_state=env.state, buffer_state=buf.init(B, tmpl),
                               metadata=decorate(env.metadata(), env.state["core"]))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    for m in range(moves):
        pre_obs, pre_mask, pre_player = obs_fn(state.env_state).cpu().numpy(), state.metadata.action_mask.cpu().numpy(), \
            state.metadata.cur_player_id.cpu().numpy()
        state = tz.collect(gen, state, None, evaluator=ev, env_step_fn=env_step_fn, env_init_fn=None, max_steps=4,
                           memory_buffer=buf, state_to_nn_input_fn=obs_fn, transform_fns=[flip], leaf_fn=game.leaf_fn,
                           dirichlet_noise=torch.distributions.Dirichlet(torch.full((B, F), 0.3)).sample().cuda())
        exps, rew, term, trunc = records[-1]
        assert len(exps) == 2
        assert np.array_equal(exps[0]["observation_nn"], pre_obs) and np.array_equal(exps[0]["policy_mask"], pre_mask)
        assert np.array_equal(exps[0]["cur_player_id"], pre_player) and not exps[0]["reward"].any()
        assert np.array_equal(exps[1]["policy_mask"], pre_mask[:, ::-1]) and np.array_equal(exps[1]["policy_weights"], exps[0]["policy_weights"][:, ::-1])
        assert np.allclose(exps[0]["policy_weights"].sum(1), 1.0, atol=1e-5)
    tmpl_np = {"reward": np.zeros((2,), np.float32), "policy_weights": np.zeros((F,), np.float32), "policy_mask": np.zeros((F,), bool),
               "observation_nn": np.zeros((4,), np.float32), "cur_player_id": np.zeros((), np.int32)}
    s = RN.init(B, cap, tmpl_np)
    saw_term = saw_trunc = False
    for exps, rew, term, trunc in records:
        RN.collect_update(s, exps, rew, term, trunc, cap)
        saw_term |= bool(term.any())
        saw_trunc |= bool(trunc.any())
    assert saw_term and saw_trunc
    got = _to_np(state.buffer_state)
    assert np.array_equal(got["next_idx"], s.next_idx) and np.array_equal(got["populated"], s.populated)
    assert np.array_equal(got["has_reward"], s.has_reward) and np.array_equal(got["episode_start_idx"], s.episode_start_idx)
    for f in FIELDS:
        assert np.array_equal(got["buf_" + f], s.buffer[f]), f
