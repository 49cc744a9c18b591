"""The replay-buffer oracle (oracle/replay_numpy.py) against fixtures produced by the reference's own
core/memory/replay_memory.py (tests/golden/make_golden_replay.py), plus hand-derived known answers."""
import os

import numpy as np
import pytest

from oracle import replay_numpy as RN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("reward", "policy_weights", "policy_mask", "observation_nn", "cur_player_id")
CASES = ["replay_small", "replay_transforms_wrap", "replay_single_player"]


def replay_fixture(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def run_oracle(fx):
    x = {k[3:]: fx[k] for k in fx.files if k.startswith("in_")}
    steps, n_exp, B = x["cur_player_id"].shape
    cap = fx["ref_populated"].shape[1]
    P, F = x["rewards"].shape[2], x["policy_weights"].shape[3]
    tmpl = {"reward": np.zeros((P,), np.float32), "policy_weights": np.zeros((F,), np.float32),
            "policy_mask": np.zeros((F,), bool), "observation_nn": np.zeros(x["observation_nn"].shape[3:], np.float32),
            "cur_player_id": np.zeros((), np.int32)}
    s = RN.init(B, cap, tmpl)
    for t in range(steps):
        exps = [{"observation_nn": x["observation_nn"][t, e], "policy_mask": x["policy_mask"][t, e],
                 "policy_weights": x["policy_weights"][t, e], "reward": np.zeros((B, P), np.float32),
                 "cur_player_id": x["cur_player_id"][t, e]} for e in range(n_exp)]
        RN.collect_update(s, exps, x["rewards"][t], x["terminated"][t], x["truncated"][t], cap)
    return s, x


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_fixture(name):
    fx = replay_fixture(name)
    s, x = run_oracle(fx)
    assert np.array_equal(s.next_idx, fx["ref_next_idx"])
    assert np.array_equal(s.episode_start_idx, fx["ref_episode_start_idx"])
    assert np.array_equal(s.populated, fx["ref_populated"])
    assert np.array_equal(s.has_reward, fx["ref_has_reward"])
    for f in FIELDS:
        assert np.array_equal(s.buffer[f], fx["ref_buf_" + f]), f
    S = fx["ref_sample_reward"].shape[0]
    assert int((s.populated & s.has_reward).sum()) >= S  # otherwise the draw would reach masked-out slots
    smp = RN.sample(s, x["gumbel"], S)
    for f in FIELDS:
        assert np.array_equal(smp[f], fx["ref_sample_" + f]), f


def test_known_answer_episode_lifecycle():
    """capacity 4, one env: two steps, terminate -> both rows rewarded; two more steps, truncate -> rolled back."""
    tmpl = {"reward": np.zeros((2,), np.float32), "x": np.zeros((), np.int32)}
    s = RN.init(1, 4, tmpl)
    for v in (10, 11):
        RN.add_experience(s, {"reward": np.zeros((1, 2), np.float32), "x": np.array([v], np.int32)}, 4)
    assert s.next_idx.tolist() == [2] and s.populated[0].tolist() == [True, True, False, False]
    assert s.has_reward[0].tolist() == [False, False, True, True]
    RN.assign_rewards(s, np.array([[1.0, -1.0]], np.float32), np.array([True]))
    assert s.buffer["reward"][0].tolist() == [[1, -1], [1, -1], [0, 0], [0, 0]] and s.has_reward.all()
    assert s.episode_start_idx.tolist() == [2]
    for v in (12, 13):
        RN.add_experience(s, {"reward": np.zeros((1, 2), np.float32), "x": np.array([v], np.int32)}, 4)
    assert s.next_idx.tolist() == [0]  # wrapped
    RN.truncate(s, np.array([True]))
    assert s.next_idx.tolist() == [2] and s.populated[0].tolist() == [True, True, False, False] and s.has_reward.all()
    assert s.buffer["x"][0].tolist() == [10, 11, 12, 13]  # truncated rows are not zeroed (replay_memory.py:123-127)


def test_masks_leave_other_envs_untouched():
    tmpl = {"reward": np.zeros((1,), np.float32)}
    s = RN.init(3, 3, tmpl)
    RN.add_experience(s, {"reward": np.zeros((3, 1), np.float32)}, 3)
    before = s.copy()
    RN.assign_rewards(s, np.ones((3, 1), np.float32), np.array([False, True, False]))
    assert s.has_reward[1].all() and not s.has_reward[0, 0] and not s.has_reward[2, 0]
    assert s.buffer["reward"][1, 0, 0] == 1 and s.buffer["reward"][0, 0, 0] == 0
    assert s.episode_start_idx.tolist() == [0, 1, 0]
    RN.truncate(s, np.array([True, False, False]))
    assert not s.populated[0].any() and s.populated[2, 0] and s.next_idx.tolist() == [0, 1, 1]
    assert np.array_equal(before.buffer["reward"][2], s.buffer["reward"][2])


def test_fixtures_regenerate_from_the_reference_source():
    """With /root/reference mounted: running the reference's replay_memory.py / common.py two_player_game_step again (on
    oracle/jaxshim) reproduces the committed replay_*.npz and two_player_*.npz byte for byte."""
    import subprocess
    import sys

    from oracle import ref_via_shim as RV

    if not RV.available():
        pytest.skip("/root/reference is not mounted here")
    for script in ("make_golden_replay.py", "make_golden_two_player.py"):
        res = subprocess.run([sys.executable, os.path.join(GOLDEN, script), "--check"], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert res.stdout.count("matches") >= 2
