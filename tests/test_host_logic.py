"""Host-side mirror of the reference API (no GPU): constructor surface, config dicts, selector registry,
class factory, partitioning -- compared with what the reference's classes expose (file:line cited per test)."""
import pytest
import torch

import turbozero_b200 as tz
from turbozero_b200 import _abi


def make(**kw):
    base = dict(eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=7, max_nodes=64, num_iterations=16)
    base.update(kw)
    return tz.MCTS(**base)


def test_mcts_constructor_defaults_and_config():  # mcts.py:19-68
    ev = make()
    assert (ev.discount, ev.temperature, ev.tiebreak_noise, ev.persist_tree) == (-1.0, 1.0, 1e-8, True)
    cfg = ev.get_config()
    assert set(cfg) == {"eval_fn", "num_iterations", "branching_factor", "max_nodes", "action_selection_config", "discount",
                        "temperature", "tiebreak_noise", "persist_tree"}
    assert cfg["action_selection_config"] == {"c": 1.0, "q_transform": "normalize_q_values", "epsilon": 1e-8}  # action_selection.py:82-88


def test_selector_registry():  # action_selection.py:35-177
    p = tz.PUCTSelector(c=2.5, epsilon=1e-6).kernel_params()
    assert p == dict(selector=_abi.TZ_SEL_PUCT, c=2.5, c1=0.0, c2=1.0, epsilon=1e-6, q_transform=_abi.TZ_QT_NORMALIZE)
    m = tz.MuZeroPUCTSelector()
    assert m.get_config() == {"c1": 1.25, "c2": 19652, "q_transform": "normalize_q_values", "epsilon": 1e-8}
    assert m.kernel_params()["selector"] == _abi.TZ_SEL_MUZERO_PUCT

    class Custom(tz.MCTSActionSelector):
        def __call__(self, tree, index, discount):
            return 0

    with pytest.raises(NotImplementedError, match="no device implementation"):
        make(action_selector=Custom())  # no CPU fallback: arbitrary Python selectors are refused at construction
    with pytest.raises(NotImplementedError, match="register_q_transform"):
        tz.PUCTSelector(q_transform=lambda *a: a[0])  # an unregistered Python callable cannot run in the kernel


def test_q_transform_registry_and_host_functions():  # action_selection.py:10-32, :70, :128
    # the registered transforms are selectable by function object or by name and land in TzSearchCfg.q_transform
    assert tz.PUCTSelector(q_transform=tz.identity_q_values).kernel_params()["q_transform"] == _abi.TZ_QT_IDENTITY
    assert tz.MuZeroPUCTSelector(q_transform="identity").kernel_params()["q_transform"] == _abi.TZ_QT_IDENTITY
    assert tz.PUCTSelector(q_transform="identity").get_config()["q_transform"] == "identity"
    assert make(action_selector=tz.PUCTSelector(q_transform="identity"))._cfg().q_transform == _abi.TZ_QT_IDENTITY
    assert make()._cfg().q_transform == _abi.TZ_QT_NORMALIZE
    # normalize_q_values is host-callable like the reference's; SURVEY.md 8c KA-2: dq = [0.4, -0, -0], only child 0 visited
    q = tz.normalize_q_values(torch.tensor([0.4, -0.0, -0.0]), torch.tensor([2, 0, 0]), 0.2, 1e-8)
    assert torch.equal(q, torch.tensor([1.0, 0.0, 0.0]))
    # batched, and identical to the NumPy oracle's restatement of the same lines
    import numpy as np
    from oracle import mcts_numpy as M

    rng = np.random.default_rng(0)
    qv = rng.standard_normal((5, 9)).astype(np.float32)
    nv = rng.integers(0, 3, (5, 9)).astype(np.int32)
    pq = rng.standard_normal((5,)).astype(np.float32)
    got = tz.normalize_q_values(torch.from_numpy(qv), torch.from_numpy(nv), torch.from_numpy(pq), 1e-8).numpy()
    for i in range(5):
        assert np.array_equal(got[i], M.normalize_q_values(qv[i], nv[i], pq[i], 1e-8))
    assert torch.equal(tz.identity_q_values(torch.from_numpy(qv), None, None, 0.0), torch.from_numpy(qv))
    # a user-registered name maps to an existing device functor id
    def my_identity(q, n, parent_q, eps):
        return q
    tz.register_q_transform(my_identity, _abi.TZ_QT_IDENTITY)
    assert tz.PUCTSelector(q_transform=my_identity).kernel_params()["q_transform"] == _abi.TZ_QT_IDENTITY


def test_search_cfg_struct_from_evaluator():
    ev = make(discount=0.5, action_selector=tz.PUCTSelector(c=1.5))
    c = ev._cfg()
    assert (c.selector, c.c, c.discount, c.weighted, c.fma_backup) == (0, 1.5, 0.5, 0, 0)
    w = tz.WeightedMCTS(q_temperature=0.5, eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=3, max_nodes=8,
                        num_iterations=2)
    c = w._cfg()
    assert c.weighted == 1 and c.inv_q_temperature == 2.0
    assert w.get_config()["q_temperature"] == 0.5  # weighted_mcts.py:35-40
    w0 = tz.WeightedMCTS(q_temperature=0.0, eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=3, max_nodes=8,
                         num_iterations=2)
    assert w0._cfg().inv_q_temperature == 0.0


def test_alphazero_class_factory():  # alphazero.py:84-98
    cls = tz.AlphaZero(tz.WeightedMCTS)
    assert cls.__name__ == "AlphaZero(WeightedMCTS)" and issubclass(cls, tz.WeightedMCTS)
    ev = cls(eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=4, max_nodes=8, num_iterations=2,
             dirichlet_alpha=0.5, dirichlet_epsilon=0.1, q_temperature=2.0)
    cfg = ev.get_config()
    assert cfg["dirichlet_alpha"] == 0.5 and cfg["dirichlet_epsilon"] == 0.1 and cfg["q_temperature"] == 2.0
    assert tz.AlphaZero(tz.MCTS)(eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=4, max_nodes=8,
                                 num_iterations=2).dirichlet_alpha == 0.3  # alphazero.py:19-20


def test_partition_and_shard_slice():  # common.py:12-29, train.py:204-217
    data = {"a": torch.arange(24).reshape(8, 3), "b": (torch.arange(8),)}
    parts = tz.partition(data, 4)
    assert parts["a"].shape == (4, 2, 3) and parts["b"][0].shape == (4, 2)
    for r in range(4):
        sl = tz.shard_slice(8, r, 4)
        assert torch.equal(parts["a"][r], data["a"][sl])
    with pytest.raises(ValueError):
        tz.shard_slice(10, 0, 4)


def test_trees_refuse_cpu_devices():
    with pytest.raises(_abi.TzError, match="no CPU implementation"):
        tz.init_tree(2, 8, 3, {"x": torch.zeros(4)}, device="cpu")


def test_eval_output_and_metadata_replace():  # evaluator.py:10-19, types.py:11-25
    md = tz.StepMetadata(rewards=torch.zeros(2, 2), action_mask=torch.ones(2, 3, dtype=torch.bool), terminated=torch.zeros(2, dtype=torch.bool),
                         cur_player_id=torch.zeros(2, dtype=torch.int32), step=torch.zeros(2, dtype=torch.int32))
    md2 = md.replace(step=md.step + 1)
    assert int(md2.step[0]) == 1 and int(md.step[0]) == 0
    out = tz.EvalOutput(eval_state=None, action=torch.zeros(2), policy_weights=torch.zeros(2, 3))
    assert out.replace(eval_state=1).eval_state == 1
