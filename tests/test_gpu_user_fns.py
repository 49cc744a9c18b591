"""The path a user of the reference takes: their OWN batched `eval_fn` and `env_step_fn` (plain torch here) driven by
`AlphaZero(MCTS).evaluate` through the built-in host glue -- mask / softmax / terminal-value select of mcts.py:165-172 and the
Dirichlet root of alphazero.py:57-76 -- with no synthetic stand-in kernel involved.

Checked against the NumPy oracle (oracle/mcts_numpy.py), tree by tree: the harness records every batch the user functions
return, restates the reference's glue lines with plain torch ops on those batches (same ops, same shapes, same device =>
same bits), and feeds the rows to the oracle's traverse / expand / backpropagate / root_action / step.  Integers bit-exact,
floats ==."""
import numpy as np
import pytest

from helpers import assert_trees_equal, tree_to_numpy
from oracle import mcts_numpy as M

pytestmark = pytest.mark.gpu

B, F, N, S, MOVES = 6, 5, 40, 24, 4


class ToyGame:
    """A batched two-player toy env in plain torch: state = {"x": float32[B, F], "t": int32[B], "player": int32[B]}."""

    def __init__(self, torch, F=F, value_scale=1.0):
        self.torch = torch
        self.F = F
        self.value_scale = value_scale  # multiplies every value / reward the search sees
        self.calls = []  # (kind, tensors...) in call order

    def init(self):
        t = self.torch
        F = self.F
        g = t.Generator(device="cuda").manual_seed(3)
        # (keys in sorted order, so that insertion order == sorted order whichever the pytree flattening uses)
        return {"player": t.zeros((B,), dtype=t.int32, device="cuda"), "t": t.zeros((B,), dtype=t.int32, device="cuda"),
                "x": t.rand((B, F), device="cuda", generator=g)}

    def metadata(self, s, tz):
        t = self.torch
        term = (s["t"] >= 5) | (s["x"][:, 0] > 0.93)
        r0 = t.where(s["x"][:, 1] > 0.5, 1.0, -1.0) * self.value_scale
        return tz.StepMetadata(rewards=t.stack([r0, -r0], 1), action_mask=s["x"] > 0.12, terminated=term,
                               cur_player_id=s["player"].clone(), step=s["t"].clone())

    def make_step_fn(self, tz):
        t = self.torch
        F = self.F

        def env_step_fn(state, action):  # core/types.py:27 (batched)
            a = action.long()
            bump = t.nn.functional.one_hot(a, F).to(t.float32) * 0.37
            x = t.frac(state["x"] * 1.7 + bump + 0.11 * (a.to(t.float32).unsqueeze(1) + 1.0))
            new = {"player": 1 - state["player"], "t": state["t"] + 1, "x": x}
            md = self.metadata(new, tz)
            self.calls.append(("step", {k: v.clone() for k, v in new.items()}, md))
            return new, md

        return env_step_fn

    def eval_fn(self, env_state, params, key):  # core/types.py:31 (batched): elementwise only, no reductions
        t = self.torch
        x = env_state["x"]
        logits = t.sin(x * 3.7 + params) * 2.0
        value = t.tanh(x[:, 0] * 1.3 - x[:, 1]) * self.value_scale
        self.calls.append(("eval", logits.clone(), value.clone()))
        return logits, value


def _emb_rows(state, b):
    return [np.ascontiguousarray(state[k][b].cpu().numpy()).view(np.uint8).reshape(-1) for k in sorted(state)]


@pytest.mark.parametrize("programmatic", [False, True])
def test_user_eval_and_env_functions_through_the_builtin_glue(programmatic):
    """`programmatic=True`: TzSearchCfg.programmatic with ordinary framework kernels between the launches (the first form of
    its contract) -- same trees."""
    _run(programmatic=programmatic)


@pytest.mark.parametrize("scale", [1e-22, 3e19, 1e-30])
@pytest.mark.parametrize("shape", ["narrow", "wide_one_warp", "wide_cta", "weighted_one_warp", "weighted_cta"])
def test_extreme_magnitudes_take_the_exact_division_path(shape, scale):
    """Values of 1e-22 / 3e19 / 1e-30 put the selector's and the weighted backup's operands outside the range in which the
    straight-line division (div_core) is proven equal to div.rn: the kernels must notice and repeat the call with the
    hardware division -- same bits as the NumPy oracle on every path (one lane per level, one warp per tree, CTA per tree;
    plain and weighted backups)."""
    _run(F=5 if shape == "narrow" else 40, value_scale=scale, weighted=shape.startswith("weighted"),
         sim_warps=4 if shape.endswith("cta") else 1, moves=3)


def _run(programmatic=False, F=F, value_scale=1.0, weighted=False, sim_warps=0, moves=MOVES):
    import torch
    import turbozero_b200 as tz

    game = ToyGame(torch, F, value_scale)
    step_fn = game.make_step_fn(tz)
    base = tz.WeightedMCTS if weighted else tz.MCTS
    ev = tz.AlphaZero(base)(eval_fn=game.eval_fn, action_selector=tz.PUCTSelector(c=1.25), branching_factor=F, max_nodes=N,
                            num_iterations=S, discount=-1.0, temperature=1.0, dirichlet_alpha=0.3, dirichlet_epsilon=0.25)
    ev.programmatic_launch = programmatic
    ev.sim_warps = sim_warps
    state = game.init()
    tree = ev.init_batched(B, {k: v[0] for k, v in state.items()})
    params = torch.tensor(0.2, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(11)
    cfg = M.SearchCfg(selector=0, c=1.25, discount=-1.0, weighted=weighted)
    emb_bytes = [int(np.prod(v.shape[1:])) * v.element_size() for k, v in sorted(state.items())]
    ref = [M.init_tree(N, F, emb_bytes, weighted=weighted) for _ in range(B)]
    finfo = torch.finfo(torch.float32)
    md = game.metadata(state, tz)
    for m in range(moves):
        game.calls.clear()
        dn = torch._sample_dirichlet(torch.full((B, F), 0.3, device="cuda"), generator=gen)
        u01 = torch.rand((B,), device="cuda", generator=gen)
        root_state = {k: v.clone() for k, v in state.items()}
        out = ev.evaluate(gen, tree, state, md, params, step_fn, dirichlet_noise=dn, uniform01=u01)
        # ---- the reference's host glue restated on the recorded batches -------------------------------------------
        kind, root_logits, root_value = game.calls[0]
        assert kind == "eval" and len(game.calls) == 1 + 2 * S
        root_policy = torch.softmax(root_logits, dim=-1)                                        # alphazero.py:58-59
        noisy = ((1 - 0.25) * root_policy) + (0.25 * dn)                                        # :63-66
        new_logits = torch.log(torch.clamp(noisy, min=finfo.tiny))                              # :69-73
        root_pol = torch.softmax(torch.where(md.action_mask, new_logits, finfo.min), dim=-1)    # :75-76
        for b in range(B):
            M.set_root(ref[b], root_pol[b].cpu().numpy(), np.float32(root_value[b].item()), _emb_rows(root_state, b))
        for s in range(S):
            _, new_state, smd = game.calls[1 + 2 * s]
            _, logits, value = game.calls[2 + 2 * s]
            pol = torch.softmax(torch.where(smd.action_mask, logits, finfo.min), dim=-1)        # mcts.py:170-171
            player_reward = smd.rewards.gather(1, smd.cur_player_id.long().unsqueeze(1)).squeeze(1)  # :166
            val = torch.where(smd.terminated, player_reward, value)                             # :172
            pol_n, val_n, term_n = pol.cpu().numpy(), val.cpu().numpy(), smd.terminated.cpu().numpy()
            for b in range(B):
                parent, action, _ = M.traverse(ref[b], cfg)
                M.expand(ref[b], parent, action, pol_n[b], np.float32(val_n[b]), bool(term_n[b]), _emb_rows(new_state, b), cfg)
                if weighted:
                    M.weighted_backpropagate(ref[b], parent, cfg)
                else:
                    M.backpropagate(ref[b], parent, np.float32(val_n[b]), cfg)
        acts = out.action.cpu().numpy()
        pws = out.policy_weights.cpu().numpy()
        got = tree_to_numpy(tree)
        for b in range(B):
            a_ref, pw_ref, _ = M.root_action(ref[b], 1.0, None, np.float32(u01[b].item()))
            assert a_ref == acts[b] and np.array_equal(pw_ref, pws[b]), (m, b)
        want = {k: np.stack([_ref_field(t, k) for t in ref]) for k in ("next_free_idx", "parents", "edge_map", "n", "p", "q", "terminated")}
        for k, v in want.items():
            assert np.array_equal(got[k], v), (m, k)
        for i, k in enumerate(sorted(state)):
            assert np.array_equal(got[f"emb{i}"].reshape(B, N, -1), np.stack([t.emb[i] for t in ref])), (m, k)
        # ---- env step + tree step (core/common.py:82-94) ----------------------------------------------------------
        state, md = step_fn(state, out.action)
        done = md.terminated
        tree = ev.step(tree, out.action, reset_mask=done)
        for b in range(B):
            if bool(done[b]):
                M.reset(ref[b])
            else:
                M.step(ref[b], int(acts[b]), True)
    assert int(tree.next_free_idx.max()) > 1


def _ref_field(t, k):
    return {"next_free_idx": np.int32(t.next_free_idx), "parents": t.parents, "edge_map": t.edge_map, "n": t.n, "p": t.p, "q": t.q,
            "terminated": t.terminated.astype(np.uint8)}[k]
