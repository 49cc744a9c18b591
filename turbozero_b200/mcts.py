"""Batched MCTS evaluator on sm_100a kernels -- the drop-in for core/evaluators/mcts/mcts.py of the reference.

Same constructor, same `init / init_batched / evaluate / step / reset / get_value / get_config` surface and the
same plug-in points (`eval_fn`, `env_step_fn`, action selector objects).  What differs, by design:

 * everything is batched (leading axis B) instead of per-environment code under `jax.vmap`;
 * `key` is a `torch.Generator` (or None / int seed); the random numbers the reference draws with `jax.random`
   can also be passed in explicitly (`root_noise`, `uniform01`), which is how parity tests pin them;
 * the tree buffers are updated in place and the same `MCTSTree` object is returned;
 * selection, expansion, backpropagation, root-action sampling and re-rooting run as CUDA kernels through the
   C-ABI (include/tz_abi.h).  There is no PyTorch / CPU fallback: without libtz_b200.so this module raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch.utils import _pytree as pytree

from . import _abi
from .action_selection import MCTSActionSelector
from .evaluator import EvalOutput, Evaluator
from .trees import MCTSTree, Tree, init_tree, _on_device, _same_device, _stream_ptr
from .types import EnvStepFn, EvalFn, StepMetadata


@dataclass(frozen=True)
class TraversalState:
    """state.py:37-44 (batched): `parent` (B,) int32, `action` (B,) int32"""
    parent: torch.Tensor
    action: torch.Tensor


@dataclass(frozen=True)
class MCTSOutput(EvalOutput):
    """state.py:59-66"""
    eval_state: MCTSTree
    policy_weights: torch.Tensor


# leaf_fn(parent_embedding, action) -> (new_embedding, policy (B,F) f32, value (B,) f32, terminated (B,) bool/uint8)
LeafFn = Callable[[Any, torch.Tensor], Tuple[Any, torch.Tensor, torch.Tensor, torch.Tensor]]


class _Scratch:
    """Per-tree-batch exchange buffers (TzWork): outputs of select, path ring."""

    def __init__(self, tree: Tree):
        dev, B = tree.device, tree.batch_size
        self.parent = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.action = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.path = torch.zeros((B, _abi.TZ_PATH_STRIDE), dtype=torch.int32, device=dev)
        # TzWork.path_spill: the levels of paths deeper than the 32-level ring (one warp per tree), or the whole path
        # record (a CTA per tree, TzSearchCfg.sim_warps > 1): max_nodes entries hold every possible path in both forms
        self.path_spill = torch.zeros((B, max(tree.capacity, 1), 2), dtype=torch.int32, device=dev)
        self.emb_parent = [torch.zeros((B, *shape), dtype=dt, device=dev) for shape, dt in tree.emb_leaf_shapes()]
        self.emb_parent_tree = tree.unflatten_embedding(self.emb_parent)
        self.select_only = self.work()

    def work(self, policy=None, value=None, terminated=None, emb_new: Optional[List[torch.Tensor]] = None,
             backprop_noise=None) -> _abi.TzWork:
        w = _abi.TzWork()
        w.parent, w.action, w.path = self.parent.data_ptr(), self.action.data_ptr(), self.path.data_ptr()
        w.path_spill, w.path_spill_cap = self.path_spill.data_ptr(), int(self.path_spill.shape[1])
        for k, t in enumerate(self.emb_parent):
            w.emb_parent[k] = t.data_ptr()
        if policy is not None:
            w.policy, w.value, w.terminated = policy.data_ptr(), value.data_ptr(), terminated.data_ptr()
            for k, t in enumerate(emb_new):
                w.emb_new[k] = t.data_ptr()
        w.backprop_noise = backprop_noise.data_ptr() if backprop_noise is not None else None
        return w


def _scratch(tree: Tree) -> _Scratch:
    if tree._scratch is None:
        tree._scratch = _Scratch(tree)
    return tree._scratch


_SEEDED: Dict[Tuple[int, str], torch.Generator] = {}


def _generator(key, device) -> Optional[torch.Generator]:
    """`key` as a torch.Generator on `device`: a Generator is used as is, None means the global RNG, and an int seed maps to
    ONE generator per (seed, device) that ADVANCES from draw to draw -- re-seeding a fresh generator on every draw would hand
    every simulation of a search (and every move) the same numbers."""
    if isinstance(key, torch.Generator):
        return key
    if isinstance(key, int):
        k = (key, str(device))
        gen = _SEEDED.get(k)
        if gen is None:
            gen = torch.Generator(device=device)
            gen.manual_seed(key)
            _SEEDED[k] = gen
        return gen
    return None


def _rand(key, shape, device) -> torch.Tensor:
    return torch.rand(shape, dtype=torch.float32, device=device, generator=_generator(key, device))


class MCTS(Evaluator):
    """Batched Monte Carlo Tree Search (mcts.py:12-432).  Not stateful: operates on `MCTSTree` objects."""

    weighted = False  # WeightedMCTS flips this (weighted_mcts.py)

    def __init__(self,
                 eval_fn: EvalFn,
                 action_selector: MCTSActionSelector,
                 branching_factor: int,
                 max_nodes: int,
                 num_iterations: int,
                 discount: float = -1.0,
                 temperature: float = 1.0,
                 tiebreak_noise: float = 1e-8,
                 persist_tree: bool = True):
        """Arguments as mcts.py:19-53."""
        super().__init__(discount=discount)
        self.eval_fn = eval_fn
        self.num_iterations = num_iterations
        self.branching_factor = branching_factor
        self.max_nodes = max_nodes
        self.action_selector = action_selector
        self.temperature = temperature
        self.tiebreak_noise = tiebreak_noise
        self.persist_tree = persist_tree
        self.fma_backup = False  # see DESIGN.md "FMA": XLA may contract mcts.py:322; default is separate mul/add
        # TzSearchCfg.programmatic: per-simulation launches overlap the tail of the leaf kernels that precede them
        # (programmatic dependent launch).  Legal whenever `leaf_fn` / `env_step_fn` + `eval_fn` enqueue ordinary
        # kernels (see include/tz_abi.h); off by default.
        self.programmatic_launch = False
        # TzSearchCfg.sim_warps: warps cooperating on one tree in the per-simulation kernel (0 = the library's choice)
        self.sim_warps = 0
        action_selector.kernel_params()  # raises now if the selector has no device implementation

    # ------------------------------------------------------------------------------------------------
    def get_config(self) -> Dict:
        """mcts.py:56-68"""
        return {
            "eval_fn": getattr(self.eval_fn, "__name__", type(self.eval_fn).__name__),
            "num_iterations": self.num_iterations,
            "branching_factor": self.branching_factor,
            "max_nodes": self.max_nodes,
            "action_selection_config": self.action_selector.get_config(),
            "discount": self.discount,
            "temperature": self.temperature,
            "tiebreak_noise": self.tiebreak_noise,
            "persist_tree": self.persist_tree
        }

    def _programmatic_bits(self) -> int:
        """TzSearchCfg.programmatic: True -> both bits (launch programmatically and signal dependents); an int is passed on."""
        v = self.programmatic_launch
        return (3 if v else 0) if isinstance(v, bool) else int(v)

    def _cfg(self) -> _abi.TzSearchCfg:
        kp = self.action_selector.kernel_params()
        q_temp = getattr(self, "q_temperature", 1.0)
        inv_t = float(np.float32(1.0 / q_temp)) if q_temp > 0 else 0.0
        return _abi.TzSearchCfg(selector=kp["selector"], c=kp["c"], c1=kp["c1"], c2=kp["c2"], epsilon=kp["epsilon"],
                                discount=self.discount, weighted=int(self.weighted), inv_q_temperature=inv_t,
                                fma_backup=int(self.fma_backup), programmatic=self._programmatic_bits(),
                                q_transform=int(kp.get("q_transform", 0)), sim_warps=int(self.sim_warps))

    # ------------------------------------------------------------------------------------------------
    def init(self, template_embedding: Any, *args, device=None, **kwargs) -> MCTSTree:  # pylint: disable=arguments-differ
        """mcts.py:417-432: one empty tree (batch of 1)."""
        return self.init_batched(1, template_embedding, device=device, **kwargs)

    def init_batched(self, batch_size: int, template_embedding: Any, *args, device=None, stats: bool = False, **kwargs) -> MCTSTree:
        """evaluator.py:42-45 + mcts.py:417-432: `batch_size` empty trees, allocated directly in batched layout."""
        return init_tree(batch_size, self.max_nodes, self.branching_factor, template_embedding,
                         weighted=self.weighted, device=device, stats=stats)

    # ------------------------------------------------------------------------------------------------
    def evaluate(self,  # pylint: disable=arguments-differ
                 key,
                 eval_state: MCTSTree,
                 env_state: Any,
                 root_metadata: StepMetadata,
                 params: Any,
                 env_step_fn: Optional[EnvStepFn],
                 *,
                 leaf_fn: Optional[LeafFn] = None,
                 root_noise: Optional[torch.Tensor] = None,
                 uniform01: Optional[torch.Tensor] = None,
                 backprop_noise: Optional[torch.Tensor] = None,
                 **kwargs) -> MCTSOutput:
        """mcts.py:71-108: populate the root, run `num_iterations` simulations, sample the root action.

        Extras over the reference signature (all optional):
        - `leaf_fn`: a fused replacement for env_step_fn + eval_fn + the mask/softmax/terminal-value glue of
          mcts.py:165-172, returning (new_embedding, policy, value, terminated) directly.
        - `root_noise` (B,F) / `uniform01` (B,): the draws `sample_root_action` would take from `key`.
        - `backprop_noise` (S,B,F): WeightedMCTS with q_temperature == 0 only (weighted_mcts.py:123).
        """
        with _on_device(eval_state.device):
            tree = self.update_root(key, eval_state, env_state, params, root_metadata=root_metadata, **kwargs)
            lib, cfg, ts = _abi.lib(), self._cfg(), tree.struct()
            sc = _scratch(tree)
            stream = _stream_ptr()
            S = self.num_iterations
            if S > 0:
                _abi.check(lib.tz_select(C.byref(ts), C.byref(cfg), C.byref(sc.select_only), stream), "tz_select")
            for s in range(S):
                bpn = self._backprop_noise(key, tree, backprop_noise, s)
                w, keep = self._leaf_work(key, tree, sc, params, env_step_fn, leaf_fn, bpn)
                fn = lib.tz_expand_backprop_select if s + 1 < S else lib.tz_expand_backprop
                _abi.check(fn(C.byref(ts), C.byref(cfg), C.byref(w), stream), "tz_expand_backprop")
                del keep
            action, policy_weights = self.sample_root_action(key, tree, root_noise=root_noise, uniform01=uniform01)
        return MCTSOutput(eval_state=tree, action=action, policy_weights=policy_weights)

    def _backprop_noise(self, key, tree, backprop_noise, s):
        return None

    def _leaf_work(self, key, tree: Tree, sc: _Scratch, params, env_step_fn, leaf_fn, backprop_noise):
        """mcts.py:160-172: the host framework's part of one simulation, between select and expand."""
        if leaf_fn is not None:
            new_embedding, policy, value, terminated = leaf_fn(sc.emb_parent_tree, sc.action)
        else:
            new_embedding, metadata = env_step_fn(sc.emb_parent_tree, sc.action)
            player_reward = torch.gather(metadata.rewards, 1, metadata.cur_player_id.long().unsqueeze(1)).squeeze(1)
            policy_logits, value = self.eval_fn(new_embedding, params, key)
            policy_logits = torch.where(metadata.action_mask.bool(), policy_logits,
                                        torch.finfo(policy_logits.dtype).min)
            policy = torch.softmax(policy_logits, dim=-1)
            terminated = metadata.terminated
            value = torch.where(terminated.bool(), player_reward.to(value.dtype), value)
        B, F = tree.batch_size, tree.branching_factor
        policy = policy.to(torch.float32).reshape(B, F).contiguous()
        value = value.to(torch.float32).reshape(B).contiguous()
        terminated = terminated.reshape(B)
        terminated = (terminated if terminated.element_size() == 1 else terminated.bool()).contiguous()
        leaves = pytree.tree_leaves(new_embedding)
        want = tree.emb_leaf_shapes()
        if len(leaves) != len(want):
            raise _abi.TzError("new embedding does not match the template embedding's pytree structure")
        emb_new = [l.to(dt).reshape(B, *shape).contiguous() for l, (shape, dt) in zip(leaves, want)]
        _same_device(tree.device, policy, value, terminated, backprop_noise, *emb_new)
        keep = (policy, value, terminated, emb_new, backprop_noise)
        return sc.work(policy, value, terminated, emb_new, backprop_noise), keep

    # ------------------------------------------------------------------------------------------------
    def get_value(self, state: MCTSTree) -> torch.Tensor:
        """mcts.py:111-120"""
        return state.data.q[:, state.ROOT_INDEX]

    def update_root(self, key, tree: MCTSTree, root_embedding: Any, params: Any, **kwargs) -> MCTSTree:  # pylint: disable=unused-argument
        """mcts.py:123-142"""
        root_policy_logits, root_value = self.eval_fn(root_embedding, params, key)
        root_policy = torch.softmax(root_policy_logits, dim=-1)
        return self._set_root(tree, root_policy, root_value, root_embedding)

    def _set_root(self, tree: MCTSTree, root_policy, root_value, root_embedding) -> MCTSTree:
        """update_root_node mcts.py:363-384 (weighted_mcts.py:66-87) + Tree.set_root tree.py:135-150"""
        B, F = tree.batch_size, tree.branching_factor
        pol = root_policy.to(torch.float32).reshape(B, F).contiguous()
        val = root_value.to(torch.float32).reshape(B).contiguous()
        leaves = [l.to(dt).reshape(B, *shape).contiguous()
                  for l, (shape, dt) in zip(pytree.tree_leaves(root_embedding), tree.emb_leaf_shapes())]
        ptrs = (C.c_void_p * max(len(leaves), 1))(*[l.data_ptr() for l in leaves])
        _same_device(tree.device, pol, val, *leaves)
        with _on_device(tree.device):
            _abi.check(_abi.lib().tz_set_root(C.byref(tree.struct()), pol.data_ptr(), val.data_ptr(), ptrs, _stream_ptr()),
                       "tz_set_root")
        return tree

    def traverse(self, tree: MCTSTree) -> TraversalState:
        """mcts.py:192-228 for every tree (one launch).  Also gathers the parents' embeddings into the
        tree's scratch (mcts.py:161-164), readable as `parent_embedding(tree)`."""
        sc = _scratch(tree)
        cfg = self._cfg()
        with _on_device(tree.device):
            _abi.check(_abi.lib().tz_select(C.byref(tree.struct()), C.byref(cfg), C.byref(sc.select_only), _stream_ptr()),
                       "tz_select")
        return TraversalState(parent=sc.parent, action=sc.action)

    @staticmethod
    def parent_embedding(tree: MCTSTree) -> Any:
        return _scratch(tree).emb_parent_tree

    def iterate(self, key, tree: MCTSTree, params: Any, env_step_fn: Optional[EnvStepFn], *, leaf_fn=None,
                backprop_noise=None) -> MCTSTree:
        """mcts.py:145-189: one un-fused simulation (select launch, host leaf evaluation, expand+backprop launch)."""
        self.traverse(tree)
        sc = _scratch(tree)
        cfg = self._cfg()
        with _on_device(tree.device):
            w, keep = self._leaf_work(key, tree, sc, params, env_step_fn, leaf_fn, backprop_noise)
            _abi.check(_abi.lib().tz_expand_backprop(C.byref(tree.struct()), C.byref(cfg), C.byref(w), _stream_ptr()),
                       "tz_expand_backprop")
        del keep
        return tree

    def sample_root_action(self, key, tree: MCTSTree, *, root_noise=None, uniform01=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """mcts.py:265-296: (action (B,) int32, policy_weights (B,F) float32)."""
        B, F, dev = tree.batch_size, tree.branching_factor, tree.device
        action = torch.empty((B,), dtype=torch.int32, device=dev)
        pw = torch.empty((B, F), dtype=torch.float32, device=dev)
        noise_ptr = u_ptr = None
        if self.temperature == 0:
            if root_noise is None:
                root_noise = _rand(key, (B, F), dev) * self.tiebreak_noise  # mcts.py:285
            root_noise = root_noise.to(torch.float32).contiguous()
            noise_ptr = root_noise.data_ptr()
        else:
            if uniform01 is None:
                uniform01 = _rand(key, (B,), dev)  # consumed by jax.random.choice, mcts.py:294
            uniform01 = uniform01.to(torch.float32).contiguous()
            u_ptr = uniform01.data_ptr()
        _same_device(dev, root_noise if self.temperature == 0 else uniform01)
        with _on_device(dev):
            _abi.check(_abi.lib().tz_root_action(C.byref(tree.struct()), float(self.temperature), noise_ptr, u_ptr, None,
                                                 pw.data_ptr(), None, action.data_ptr(), _stream_ptr()), "tz_root_action")
        return action, pw

    def root_visits(self, tree: MCTSTree) -> torch.Tensor:
        """tree.get_child_data('n', ROOT) (mcts.py:276) as one kernel: (B,F) int32."""
        visits = torch.empty((tree.batch_size, tree.branching_factor), dtype=torch.int32, device=tree.device)
        with _on_device(tree.device):
            _abi.check(_abi.lib().tz_root_action(C.byref(tree.struct()), 1.0, None, None, visits.data_ptr(), None, None, None,
                                                 _stream_ptr()), "tz_root_action")
        return visits

    # ------------------------------------------------------------------------------------------------
    def reset(self, state: MCTSTree, mask: Optional[torch.Tensor] = None) -> MCTSTree:
        """mcts.py:387-396.  `mask` (B,) restricts the reset to the flagged trees (others untouched)."""
        flags = None
        if mask is not None:
            flags = torch.where(mask.bool(), 1, 2).to(torch.uint8).contiguous()  # 1 = reset, 2 = leave untouched
            _same_device(state.device, flags)
        # persist_tree = 0: no tree is re-rooted here, so no action is needed (flag 2 returns before the reset test)
        with _on_device(state.device):
            _abi.check(_abi.lib().tz_reroot(C.byref(state.struct()), None, None if flags is None else flags.data_ptr(), 0,
                                            _stream_ptr()), "tz_reroot")
        return state

    def step(self, state: MCTSTree, action: torch.Tensor, reset_mask: Optional[torch.Tensor] = None,
             keep_mask: Optional[torch.Tensor] = None) -> MCTSTree:
        """mcts.py:399-414.  `reset_mask` folds the caller's reset-vs-step select (core/common.py:89-94) into the
        same launch: flagged trees are reset instead of re-rooted; `keep_mask` trees are left untouched."""
        flags = None
        if reset_mask is not None and keep_mask is None:
            # the common case (core/common.py:89-94 with reset=True): a bool / uint8 mask IS the flag byte (1 = reset)
            rm = reset_mask
            flags = rm.view(torch.uint8) if rm.dtype == torch.bool else (rm != 0).view(torch.uint8)
            flags = flags.reshape(state.batch_size).contiguous()
        elif reset_mask is not None or keep_mask is not None:
            flags = torch.zeros((state.batch_size,), dtype=torch.uint8, device=state.device)
            if reset_mask is not None:
                flags = torch.where(reset_mask.bool(), 1, flags.int()).to(torch.uint8)
            if keep_mask is not None:
                flags = torch.where(keep_mask.bool(), 2, flags.int()).to(torch.uint8)
            flags = flags.contiguous()
        act = action.to(torch.int32).contiguous()
        _same_device(state.device, act, flags)
        with _on_device(state.device):
            _abi.check(_abi.lib().tz_reroot(C.byref(state.struct()), act.data_ptr(), None if flags is None else flags.data_ptr(),
                                            1 if self.persist_tree else 0, _stream_ptr()), "tz_reroot")
        return state
