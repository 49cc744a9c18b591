"""Device-side synthetic game (libtz_synth.so, standin/include/tz_synth.h) wrapped for PyTorch.

Stand-in for the user's pgx environment + network, which cannot run in this image.  Used by bench.py, the GPU
parity tests and smoke(); it is not part of the product path (the search never depends on it).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from turbozero_b200 import _abi
from turbozero_b200.trees import Tree, _stream_ptr

from . import abi as _sabi

# name: (F, payload_bytes, rho256, tau1024, max_depth): shapes of BASELINE.json's pgx games
GAMES = {
    "tic_tac_toe": (9, 72, 154, 40, 9),
    "connect_four": (7, 272, 230, 12, 42),
    "othello": (65, 448, 38, 6, 60),
    "go_9x9": (82, 4080, 205, 2, 120),
    "2048": (4, 560, 218, 4, 200),
}


@dataclass
class SyntheticGame:
    F: int
    payload_bytes: int
    rho256: int
    tau1024: int
    max_depth: int
    seed: int

    @classmethod
    def named(cls, name: str, seed: int) -> "SyntheticGame":
        F, P, rho, tau, D = GAMES[name]
        return cls(F, P, rho, tau, D, seed)

    def __post_init__(self):
        self._c = _sabi.TzSynthGame(F=self.F, payload_bytes=self.payload_bytes, rho256=self.rho256, tau1024=self.tau1024,
                                   max_depth=self.max_depth, seed=self.seed)
        self._out: Dict[int, tuple] = {}

    @property
    def emb_bytes(self) -> int:
        return 16 + self.payload_bytes

    def template_embedding(self) -> Dict[str, torch.Tensor]:
        t = {"core": torch.zeros(4, dtype=torch.int32)}
        if self.payload_bytes > 0:
            t["payload"] = torch.zeros(self.payload_bytes, dtype=torch.uint8)
        return t

    # --- env_init_fn ---
    def init_states(self, B: int, env_offset: int = 0, episode: Optional[torch.Tensor] = None, device="cuda"):
        episode = torch.zeros((B,), dtype=torch.int32, device=device) if episode is None else episode
        state = {"core": torch.empty((B, 4), dtype=torch.int32, device=device)}
        if self.payload_bytes > 0:
            state["payload"] = torch.empty((B, self.payload_bytes), dtype=torch.uint8, device=device)
        _abi.check(_sabi.synth_lib().tz_synth_init_states(C.byref(self._c), B, env_offset, episode.data_ptr(),
                                                         state["core"].data_ptr(), self._pay(state), _stream_ptr()),
                   "tz_synth_init_states")
        return state, episode

    @staticmethod
    def _pay(state) -> Optional[int]:
        return state["payload"].data_ptr() if "payload" in state else None

    # --- root evaluation (mcts.py:137-138 / alphazero.py:57-76) ---
    def root_eval(self, state, dir_noise: Optional[torch.Tensor] = None, dir_eps: float = 0.25,
                  out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        B, dev = state["core"].shape[0], state["core"].device
        pol, val = out if out is not None else (torch.empty((B, self.F), dtype=torch.float32, device=dev),
                                                torch.empty((B,), dtype=torch.float32, device=dev))
        _abi.check(_sabi.synth_lib().tz_synth_root(C.byref(self._c), B, state["core"].data_ptr(),
                                                  None if dir_noise is None else dir_noise.data_ptr(), dir_eps,
                                                  pol.data_ptr(), val.data_ptr(), _stream_ptr()), "tz_synth_root")
        return pol, val

    # --- fused leaf_fn (env_step_fn + eval_fn + mcts.py:166-172) ---
    def leaf_fn(self, parent_emb, action: torch.Tensor):
        B, dev = action.shape[0], action.device
        bufs = self._out.get(B)
        if bufs is None:
            new = {"core": torch.empty((B, 4), dtype=torch.int32, device=dev)}
            if self.payload_bytes > 0:
                new["payload"] = torch.empty((B, self.payload_bytes), dtype=torch.uint8, device=dev)
            bufs = (new, torch.empty((B, self.F), dtype=torch.float32, device=dev),
                    torch.empty((B,), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.uint8, device=dev))
            self._out[B] = bufs
        new, pol, val, term = bufs
        _abi.check(_sabi.synth_lib().tz_synth_leaf(C.byref(self._c), B, parent_emb["core"].data_ptr(), action.data_ptr(),
                                                  pol.data_ptr(), val.data_ptr(), term.data_ptr(), new["core"].data_ptr(),
                                                  self._pay(new), _stream_ptr()), "tz_synth_leaf")
        return new, pol, val, term

    # --- real environment step after a move (common.py:82-99) ---
    def env_step(self, state, action: torch.Tensor, episode: torch.Tensor, reset_flag: torch.Tensor, env_offset: int = 0):
        B = action.shape[0]
        _abi.check(_sabi.synth_lib().tz_synth_env_step(C.byref(self._c), B, env_offset, action.data_ptr(),
                                                      state["core"].data_ptr(), self._pay(state), episode.data_ptr(),
                                                      reset_flag.data_ptr(), _stream_ptr()), "tz_synth_env_step")
        return state

    def leaf_callback(self, B: int):
        """(function pointer, user pointer, keepalive) for tz_search: the whole simulation loop stays in C."""
        ctx = _sabi.TzSynthCtx(game=self._c, B=B)
        fn = C.cast(_sabi.synth_lib().tz_synth_leaf_cb, C.c_void_p)
        return fn, C.cast(C.pointer(ctx), C.c_void_p), ctx


class _SelfPlayLane:
    """Envs [b0, b1) of a SyntheticSelfPlay: views into its buffers plus the structs the C-ABI takes."""

    def __init__(self, sp: "SyntheticSelfPlay", b0: int, b1: int, stream: Optional[torch.cuda.Stream]):
        self.sp, self.b0, self.b1, self.stream = sp, b0, b1, stream
        sl = slice(b0, b1)
        self.tree = sp.tree if (b0 == 0 and b1 == sp.B) else sp.tree.slice(b0, b1)
        self.state = {k: v[sl] for k, v in sp.state.items()}
        self.episode, self.reset_flag = sp.episode[sl], sp.reset_flag[sl]
        self.root_policy, self.root_value = sp.root_policy[sl], sp.root_value[sl]
        self.dir_noise = None if sp.dir_noise is None else sp.dir_noise[sl]
        self.root_noise, self.uniform01 = sp.root_noise[sl], sp.uniform01[sl]
        self.action, self.policy_weights = sp.action[sl], sp.policy_weights[sl]
        w = _abi.TzWork()
        w.parent, w.action = sp.w_parent[sl].data_ptr(), sp.w_action[sl].data_ptr()
        w.path = sp.w_path[sl].data_ptr() if sp.w_path is not None else None
        if sp.w_path_spill is not None:
            w.path_spill, w.path_spill_cap = sp.w_path_spill[sl].data_ptr(), int(sp.w_path_spill.shape[1])
        w.policy, w.value, w.terminated = sp.w_policy[sl].data_ptr(), sp.w_value[sl].data_ptr(), sp.w_term[sl].data_ptr()
        for k in range(len(sp.w_emb_parent)):
            w.emb_parent[k] = sp.w_emb_parent[k][sl].data_ptr()
            w.emb_new[k] = sp.w_emb_new[k][sl].data_ptr()
        self.work = w
        self.cb = sp.game.leaf_callback(b1 - b0)
        self.done = torch.cuda.Event() if stream is not None else None

    def move(self) -> None:
        sp = self.sp
        lib, ev, game, stream = _abi.lib(), sp.ev, sp.game, _stream_ptr()
        ts = self.tree.struct()
        game.root_eval(self.state, self.dir_noise, sp.dir_eps, out=(self.root_policy, self.root_value))
        ptrs = (C.c_void_p * 2)(self.state["core"].data_ptr(), SyntheticGame._pay(self.state))
        _abi.check(lib.tz_set_root(C.byref(ts), self.root_policy.data_ptr(), self.root_value.data_ptr(), ptrs, stream), "tz_set_root")
        fn, user, _keep = self.cb
        _abi.check(lib.tz_search(C.byref(ts), C.byref(sp.cfg), C.byref(self.work), ev.num_iterations, fn, user, stream), "tz_search")
        _abi.check(lib.tz_root_action(C.byref(ts), float(ev.temperature), self.root_noise.data_ptr(), self.uniform01.data_ptr(),
                                      None, self.policy_weights.data_ptr(), None, self.action.data_ptr(), stream), "tz_root_action")
        game.env_step(self.state, self.action, self.episode, self.reset_flag, sp.env_offset + self.b0)
        evs = sp.reroot_events  # optional (start, end) CUDA events around the re-root launch (bench instrumentation)
        if evs is not None:
            evs[0].record()
        _abi.check(lib.tz_reroot(C.byref(ts), self.action.data_ptr(), self.reset_flag.data_ptr(), 1 if ev.persist_tree else 0,
                                 stream), "tz_reroot")
        if evs is not None:
            evs[1].record()


class SyntheticSelfPlay:
    """Self-play of B synthetic games through the C-ABI only (no per-simulation Python): per move
    tz_synth_root -> tz_set_root -> tz_search(tz_synth_leaf_cb) -> tz_root_action -> tz_synth_env_step -> tz_reroot.
    This is `step_env_and_evaluator` (core/common.py:32-103) with the user's functions replaced by the stand-in.
    All buffers are static, so one move is capturable in a CUDA graph.

    `pipelines = K > 1` splits the env batch into K contiguous slices whose moves run CONCURRENTLY on K streams
    (fork / join around the move): trees never interact, and at ~1 K trees one dependent chain of small launches
    leaves most of a B200 idle, so independent chains overlap almost for free.  Results are identical per tree."""

    def __init__(self, game: SyntheticGame, evaluator, B: int, *, env_offset: int = 0, dirichlet: bool = True,
                 device="cuda", stats: bool = True, use_path: bool = True, pipelines: int = 1,
                 use_spill: bool = True):
        self.game, self.ev, self.B, self.env_offset = game, evaluator, B, env_offset
        self.dev = torch.device(device)
        self.tree: Tree = evaluator.init_batched(B, game.template_embedding(), device=device, stats=stats)
        self.state, self.episode = game.init_states(B, env_offset, device=device)
        F = game.F
        f32, dev = torch.float32, self.dev
        self.root_policy = torch.empty((B, F), dtype=f32, device=dev)
        self.root_value = torch.empty((B,), dtype=f32, device=dev)
        self.dir_noise = torch.empty((B, F), dtype=f32, device=dev) if dirichlet else None
        self.root_noise = torch.zeros((B, F), dtype=f32, device=dev)
        self.uniform01 = torch.zeros((B,), dtype=f32, device=dev)
        self.action = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.policy_weights = torch.zeros((B, F), dtype=f32, device=dev)
        self.reset_flag = torch.zeros((B,), dtype=torch.uint8, device=dev)
        # TzWork with static leaf buffers
        self.w_parent = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.w_action = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.w_path = torch.zeros((B, _abi.TZ_PATH_STRIDE), dtype=torch.int32, device=dev) if use_path else None
        N_ = evaluator.max_nodes
        self.w_path_spill = (torch.zeros((B, max(N_, 1), 2), dtype=torch.int32, device=dev)
                             if (use_path and use_spill) else None)
        self.w_policy = torch.empty((B, F), dtype=f32, device=dev)
        self.w_value = torch.empty((B,), dtype=f32, device=dev)
        self.w_term = torch.empty((B,), dtype=torch.uint8, device=dev)
        leaves = [torch.empty((B, 4), dtype=torch.int32, device=dev)]
        leaves2 = [torch.empty((B, 4), dtype=torch.int32, device=dev)]
        if game.payload_bytes > 0:
            leaves.append(torch.empty((B, game.payload_bytes), dtype=torch.uint8, device=dev))
            leaves2.append(torch.empty((B, game.payload_bytes), dtype=torch.uint8, device=dev))
        self.w_emb_parent, self.w_emb_new = leaves, leaves2
        self.cfg = evaluator._cfg()
        self.reroot_events = None
        self.dir_eps = getattr(evaluator, "dirichlet_epsilon", 0.25)
        K = max(1, min(int(pipelines), B))
        bounds = [(k * B) // K for k in range(K + 1)]
        self.lanes = [_SelfPlayLane(self, bounds[k], bounds[k + 1], torch.cuda.Stream(device=dev) if K > 1 else None)
                      for k in range(K)]
        self._fork = torch.cuda.Event() if K > 1 else None
        # single-lane aliases kept for callers that drive the C-ABI themselves
        self.work, self._cb = self.lanes[0].work, self.lanes[0].cb

    @property
    def pipelines(self) -> int:
        return len(self.lanes)

    # --- measurement: in-step launch timeline (TzWork.timeline + tz_synth_set_timeline), single lane only ---
    TL_SLOTS = 2048

    def timeline_begin(self) -> None:
        """From now on every per-simulation search launch and every leaf launch issued (or captured) by this object
        records {first warp in, last warp has its inputs, last warp out} in %globaltimer ns."""
        assert len(self.lanes) == 1
        init = torch.zeros((self.TL_SLOTS, 4), dtype=torch.int64, device=self.dev)
        init[:, 0] = -1  # ~0 as uint64
        self._tl_init = init
        self._tl_search, self._tl_leaf = init.clone(), init.clone()
        self.lanes[0].work.timeline = self._tl_search.data_ptr()
        self.lanes[0].work.timeline_slots = self.TL_SLOTS
        _abi.check(_sabi.synth_lib().tz_synth_set_timeline(self._tl_leaf.data_ptr(), self.TL_SLOTS), "tz_synth_set_timeline")

    def timeline_mark(self):
        """(search seq, leaf seq) of the next launches: call right before issuing / capturing the move to be read."""
        return int(_abi.lib().tz_launch_seq()), int(_sabi.synth_lib().tz_synth_leaf_seq())

    def timeline_clear(self) -> None:
        self._tl_search.copy_(self._tl_init)
        self._tl_leaf.copy_(self._tl_init)

    def timeline_read(self, mark, n_search: int, n_leaf: int):
        """Rows of the `n_search` search launches / `n_leaf` leaf launches issued after `mark`, as int64 numpy [n,3]."""
        s0, l0 = mark
        idx_s = (torch.arange(n_search, device=self.dev) + s0) % self.TL_SLOTS
        idx_l = (torch.arange(n_leaf, device=self.dev) + l0) % self.TL_SLOTS
        return self._tl_search[idx_s, :3].cpu().numpy(), self._tl_leaf[idx_l, :3].cpu().numpy()

    def timeline_end(self) -> None:
        self.lanes[0].work.timeline = None
        self.lanes[0].work.timeline_slots = 0
        _abi.check(_sabi.synth_lib().tz_synth_set_timeline(None, 0), "tz_synth_set_timeline")

    def launches_per_move(self) -> int:
        S = self.ev.num_iterations
        # root, set_root, select + S leaf + S expand, root_action, env_step, reroot -- per lane
        return len(self.lanes) * (1 + 1 + (1 + S + S) + 1 + 1 + 1)

    def move(self) -> None:
        """One self-play move for all B games; enqueues only (no sync)."""
        if len(self.lanes) == 1:
            self.lanes[0].cb = self._cb  # honours a caller-installed leaf callback (bench instrumentation)
            self.lanes[0].move()
            return
        main = torch.cuda.current_stream()
        self._fork.record(main)
        for lane in self.lanes:
            with torch.cuda.stream(lane.stream):
                lane.stream.wait_event(self._fork)
                lane.move()
                lane.done.record(lane.stream)
        for lane in self.lanes:
            main.wait_event(lane.done)


def make_synthetic_evaluator(base, game: SyntheticGame, dir_eps: Optional[float] = 0.25, fma_backup: bool = False,
                             programmatic: bool = False, **kwargs):
    """An evaluator of class `base` (MCTS / WeightedMCTS) whose root evaluation is the synthetic stand-in's
    (tz_synth_root: mcts.py:137-138, or alphazero.py:57-76 when Dirichlet noise is passed), so that root policies
    carry the same bits as the CPU oracle's.  Everything else is the product path unchanged."""

    class SyntheticRoot(base):
        def update_root(self, key, tree, root_embedding, params, root_metadata=None, dirichlet_noise=None, **kw):
            pol, val = game.root_eval(root_embedding, dirichlet_noise, dir_eps if dir_eps is not None else 0.0)
            return self._set_root(tree, pol, val, root_embedding)

    SyntheticRoot.__name__ = f"SyntheticRoot({base.__name__})"
    ev = SyntheticRoot(eval_fn=None, branching_factor=game.F, **kwargs)
    ev.fma_backup = fma_backup
    ev.programmatic_launch = programmatic
    if programmatic:  # the stand-in's leaf kernel then waits-then-signals, as TzSearchCfg.programmatic requires
        _sabi.synth_lib().tz_synth_set_programmatic(1)
    ev.dirichlet_epsilon = dir_eps
    return ev


class SyntheticEnv:
    """The synthetic game as the `env_step_fn` of core/common.py:32-103 (batched): stepping an env whose episode
    ends re-initialises it in the same launch (common.py:95-100) and reports `terminated` for that step."""

    def __init__(self, game: SyntheticGame, B: int, env_offset: int = 0, device="cuda"):
        from turbozero_b200.types import StepMetadata

        self._md = StepMetadata
        self.game, self.B, self.env_offset = game, B, env_offset
        self.state, self.episode = game.init_states(B, env_offset, device=device)
        dev = self.state["core"].device
        self.reset_flag = torch.zeros((B,), dtype=torch.uint8, device=dev)
        self._step = torch.zeros((B,), dtype=torch.int32, device=dev)
        self._rewards = torch.zeros((B, 2), dtype=torch.float32, device=dev)
        self._mask = torch.ones((B, game.F), dtype=torch.bool, device=dev)
        self._player = torch.zeros((B,), dtype=torch.int32, device=dev)

    def metadata(self):
        return self._md(rewards=self._rewards, action_mask=self._mask, terminated=self.reset_flag, cur_player_id=self._player,
                        step=self._step)

    def env_step_fn(self, state, action: torch.Tensor):
        self.game.env_step(state, action.to(torch.int32), self.episode, self.reset_flag, self.env_offset)
        return state, self.metadata()
