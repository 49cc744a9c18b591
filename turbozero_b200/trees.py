"""Batched fixed-capacity trees in device memory -- the `Tree` / `MCTSTree` pytree of the reference
(core/trees/tree.py:11-23, core/evaluators/mcts/state.py:12-34) with a leading batch axis on every leaf
(core/evaluators/evaluator.py:42-45), laid out exactly as the kernels expect (include/tz_abi.h TzTree).

The reference's Tree methods are pure functions that return new pytrees (XLA aliases the buffers).  Here the
buffers are updated in place by the kernels and the same object is handed back.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, List, Optional

import torch
from torch.utils import _pytree as pytree

from . import _abi

NULL_INDEX = -1  # tree.py:21
NULL_VALUE = 0  # tree.py:22
ROOT_INDEX = 0  # tree.py:23


@dataclass
class MCTSNode:
    """state.py:12-30 (+ `r`, weighted_mcts.py:14-17).  Leaves carry leading axes (B, N)."""
    n: torch.Tensor  # (B,N) int32 visit count
    p: torch.Tensor  # (B,N,F) float32 policy
    q: torch.Tensor  # (B,N) float32 value estimate
    terminated: torch.Tensor  # (B,N) bool
    embedding: Any  # pytree, leaves (B,N,...)
    r: Optional[torch.Tensor] = None  # (B,N) float32 raw leaf value (WeightedMCTSNode)

    @property
    def w(self) -> torch.Tensor:
        """state.py:27-30 cumulative value estimate"""
        return self.q * self.n


WeightedMCTSNode = MCTSNode


def _stream_ptr(device=None) -> int:
    """The current CUDA stream of `device` (default: the current device) as the raw cudaStream_t the C-ABI takes."""
    return torch.cuda.current_stream(device).cuda_stream


class _on_device:
    """Context for one or more C-ABI calls on the tensors of `device`: kernels launch on -- and per-device function
    attributes are set on -- the calling thread's CURRENT device, so a tree that lives elsewhere must make its device
    current first (a process that drives several GPUs, or a tree not on the current device).  Free when it already is."""
    __slots__ = ("_ctx",)

    def __init__(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self._ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            return self._ctx.__exit__(*exc)
        return False


def _same_device(device: torch.device, *tensors) -> None:
    for t in tensors:
        if t is not None and t.device != device:
            raise _abi.TzError(f"tensor on {t.device} passed to a tree batch on {device}: all inputs of a call must live "
                               "on the tree's device")


@dataclass
class Tree:
    """tree.py:11-19 with a batch axis: next_free_idx (B,), parents (B,N), edge_map (B,N,F), data leaves (B,N,...)."""
    next_free_idx: torch.Tensor
    parents: torch.Tensor
    edge_map: torch.Tensor
    data: MCTSNode
    stats: Optional[torch.Tensor] = None  # (B,4) int64 counters, see include/tz_abi.h
    child_stats: Optional[torch.Tensor] = None  # (B,N,F,4) int32 derived table (include/tz_abi.h TzTree.child_stats)
    best: Optional[torch.Tensor] = None  # (B,N,2) int32 derived table: the selector's decision per node (TzTree.best)
    sel_state: Optional[torch.Tensor] = None  # (B,8) int32 selector parameters `best` was computed with
    _emb_leaves: List[torch.Tensor] = field(default_factory=list, repr=False)
    _emb_spec: Any = field(default=None, repr=False)
    _struct: Any = field(default=None, repr=False)
    _scratch: Any = field(default=None, repr=False)

    NULL_INDEX = NULL_INDEX
    NULL_VALUE = NULL_VALUE
    ROOT_INDEX = ROOT_INDEX

    # --- shape properties (tree.py:25-34) ---
    @property
    def batch_size(self) -> int:
        return self.parents.shape[0]

    @property
    def capacity(self) -> int:
        return self.parents.shape[-1]

    @property
    def branching_factor(self) -> int:
        return self.edge_map.shape[-1]

    @property
    def device(self) -> torch.device:
        return self.parents.device

    # --- read helpers (tree.py:37-98); plain tensor indexing, not on the hot path ---
    def data_at(self, index: int) -> MCTSNode:
        """tree.py:37-49 for every tree of the batch."""
        d = self.data
        return MCTSNode(n=d.n[:, index], p=d.p[:, index], q=d.q[:, index], terminated=d.terminated[:, index],
                        embedding=pytree.tree_map(lambda x: x[:, index], d.embedding),
                        r=None if d.r is None else d.r[:, index])

    def is_edge(self, parent_index: int, edge_index: int) -> torch.Tensor:
        """tree.py:65-75"""
        return self.edge_map[:, parent_index, edge_index] != NULL_INDEX

    def get_child_data(self, x: str, index: int, null_value=None) -> torch.Tensor:
        """tree.py:78-98"""
        if null_value is None:
            null_value = NULL_VALUE
        mapping = self.edge_map[:, index].long()
        arr = getattr(self.data, x)
        child = torch.gather(arr, 1, mapping.clamp(min=0))
        return torch.where(mapping == NULL_INDEX, torch.as_tensor(null_value, dtype=arr.dtype, device=arr.device), child)

    # --- the struct the kernels take ---
    def struct(self) -> _abi.TzTree:
        if self._struct is None:
            d = self.data
            if self.child_stats is None or self.best is None or self.sel_state is None:
                raise _abi.TzError("tree has no derived tables: build trees with init_tree / Evaluator.init_batched")
            for t in (self.next_free_idx, self.parents, self.edge_map, d.n, d.p, d.q, d.terminated, self.child_stats,
                      self.best, self.sel_state, *self._emb_leaves):
                assert t.is_cuda and t.is_contiguous(), "tree leaves must be contiguous CUDA tensors"
            assert self.parents.dtype == torch.int32 and self.edge_map.dtype == torch.int32 and d.n.dtype == torch.int32
            assert d.p.dtype == torch.float32 and d.q.dtype == torch.float32 and d.terminated.element_size() == 1
            if len(self._emb_leaves) > _abi.TZ_MAX_EMB:
                raise _abi.TzError(f"embedding has {len(self._emb_leaves)} leaves; the C-ABI carries at most {_abi.TZ_MAX_EMB}")
            s = _abi.TzTree(B=self.batch_size, N=self.capacity, F=self.branching_factor, n_emb=len(self._emb_leaves))
            s.next_free_idx = self.next_free_idx.data_ptr()
            s.parents = self.parents.data_ptr()
            s.edge_map = self.edge_map.data_ptr()
            s.n, s.p, s.q = d.n.data_ptr(), d.p.data_ptr(), d.q.data_ptr()
            s.r = d.r.data_ptr() if d.r is not None else None
            s.terminated = d.terminated.data_ptr()
            s.child_stats = self.child_stats.data_ptr()
            s.best = self.best.data_ptr()
            s.sel_state = self.sel_state.data_ptr()
            for k, leaf in enumerate(self._emb_leaves):
                s.emb[k] = leaf.data_ptr()
                s.emb_row_bytes[k] = leaf[0, 0].numel() * leaf.element_size()
            s.stats = self.stats.data_ptr() if self.stats is not None else None
            self._struct = s
        return self._struct

    def emb_leaf_shapes(self):
        return [(tuple(l.shape[2:]), l.dtype) for l in self._emb_leaves]

    def unflatten_embedding(self, leaves: List[torch.Tensor]):
        return pytree.tree_unflatten(leaves, self._emb_spec)

    # --- mutating ops (in place, stream-ordered, no sync) ---
    def reset(self) -> "Tree":
        """tree.py:272-278 for the whole batch (every row rewritten)."""
        with _on_device(self.device):
            _abi.check(_abi.lib().tz_tree_init(C.byref(self.struct()), _stream_ptr()), "tz_tree_init")
        return self

    def get_subtree(self, subtree_index: torch.Tensor, reset_mask: Optional[torch.Tensor] = None) -> "Tree":
        """tree.py:220-269 per tree, with the caller's reset-vs-step select (core/common.py:89-94) folded in:
        trees whose `reset_mask` is set are reset (tree.py:272-278) instead."""
        act = subtree_index.to(torch.int32).contiguous()
        rm = None if reset_mask is None else reset_mask.to(torch.uint8).contiguous()
        _same_device(self.device, act, rm)
        with _on_device(self.device):
            _abi.check(_abi.lib().tz_reroot(C.byref(self.struct()), act.data_ptr(), None if rm is None else rm.data_ptr(), 1,
                                            _stream_ptr()), "tz_reroot")
        return self

    def rebuild_child_stats(self) -> "Tree":
        """Recomputes the derived child_stats table from edge_map / p / q / n / terminated and forgets the cached selector
        decisions (needed only after those leaves were written from outside the kernels)."""
        with _on_device(self.device):
            _abi.check(_abi.lib().tz_rebuild_child_stats(C.byref(self.struct()), _stream_ptr()), "tz_rebuild_child_stats")
        return self

    def slice(self, start: int, stop: int) -> "Tree":
        """Trees [start, stop) of this batch as a batch of their own, sharing memory (every leaf is batch-major, so a
        contiguous range of trees is itself a dense tree batch).  Independent slices can be searched concurrently on
        different CUDA streams -- the trees never interact."""
        d = self.data
        sl = lambda x: None if x is None else x[start:stop]
        leaves = [l[start:stop] for l in self._emb_leaves]
        nd = MCTSNode(n=d.n[start:stop], p=d.p[start:stop], q=d.q[start:stop], terminated=d.terminated[start:stop],
                      embedding=pytree.tree_unflatten(leaves, self._emb_spec), r=sl(d.r))
        return Tree(self.next_free_idx[start:stop], self.parents[start:stop], self.edge_map[start:stop], nd, sl(self.stats),
                    sl(self.child_stats), sl(self.best), sl(self.sel_state), leaves, self._emb_spec)

    def clone(self) -> "Tree":
        d = self.data
        leaves = [l.clone() for l in self._emb_leaves]
        nd = MCTSNode(n=d.n.clone(), p=d.p.clone(), q=d.q.clone(), terminated=d.terminated.clone(),
                      embedding=pytree.tree_unflatten(leaves, self._emb_spec), r=None if d.r is None else d.r.clone())
        return Tree(self.next_free_idx.clone(), self.parents.clone(), self.edge_map.clone(), nd,
                    None if self.stats is None else self.stats.clone(),
                    None if self.child_stats is None else self.child_stats.clone(),
                    None if self.best is None else self.best.clone(),
                    None if self.sel_state is None else self.sel_state.clone(), leaves, self._emb_spec)


MCTSTree = Tree  # state.py:34


def init_tree(batch_size: int, max_nodes: int, branching_factor: int, template_embedding: Any, *, weighted: bool = False,
              device=None, stats: bool = False) -> Tree:
    """tree.py:281-298 + mcts.py:417-432 for a batch: indices -1, data zero, next_free_idx 0.
    `template_embedding` is an UNBATCHED pytree of tensors; only shapes and dtypes are used."""
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise _abi.TzError("turbozero_b200 trees live in CUDA memory; there is no CPU implementation of the search")
    B, N, F = batch_size, max_nodes, branching_factor
    tmpl_leaves, spec = pytree.tree_flatten(template_embedding)
    leaves = [torch.zeros((B, N, *torch.as_tensor(t).shape), dtype=torch.as_tensor(t).dtype, device=dev) for t in tmpl_leaves]
    node = MCTSNode(
        n=torch.zeros((B, N), dtype=torch.int32, device=dev),
        p=torch.zeros((B, N, F), dtype=torch.float32, device=dev),
        q=torch.zeros((B, N), dtype=torch.float32, device=dev),
        terminated=torch.zeros((B, N), dtype=torch.bool, device=dev),
        embedding=pytree.tree_unflatten(leaves, spec),
        r=torch.zeros((B, N), dtype=torch.float32, device=dev) if weighted else None,
    )
    child_stats = torch.zeros((B, N, F, 4), dtype=torch.int32, device=dev)
    child_stats[..., 3] = NULL_INDEX  # null entry {0, 0, 0, -1}
    return Tree(
        next_free_idx=torch.zeros((B,), dtype=torch.int32, device=dev),
        parents=torch.full((B, N), NULL_INDEX, dtype=torch.int32, device=dev),
        edge_map=torch.full((B, N, F), NULL_INDEX, dtype=torch.int32, device=dev),
        data=node,
        stats=torch.zeros((B, 4), dtype=torch.int64, device=dev) if stats else None,
        child_stats=child_stats,
        best=torch.full((B, N, 2), -1, dtype=torch.int32, device=dev),
        sel_state=torch.zeros((B, _abi.TZ_SEL_STATE_WORDS), dtype=torch.int32, device=dev),
        _emb_leaves=leaves, _emb_spec=spec,
    )
