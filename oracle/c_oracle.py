"""ORACLE (test infrastructure) -- ctypes wrapper around oracle/tz_oracle.c operating on NumPy host arrays laid
out exactly like the device trees (include/tz_abi.h).  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from standin.abi import TzSynthGame  # struct layout only (standin/include/tz_synth.h)
from turbozero_b200._abi import TZ_MAX_EMB, TzSearchCfg, TzTree, TzWork  # struct layouts only
from . import build as _build

_lib = None
_vp = C.c_void_p
_P = C.POINTER


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_build.build()))
        sig = {
            "tzo_tree_init": [_P(TzTree)],
            "tzo_set_root": [_P(TzTree), _vp, _vp, _P(_vp)],
            "tzo_select": [_P(TzTree), _P(TzSearchCfg), _P(TzWork)],
            "tzo_expand_backprop": [_P(TzTree), _P(TzSearchCfg), _P(TzWork)],
            "tzo_root_action": [_P(TzTree), C.c_float, _vp, _vp, _vp, _vp, _vp, _vp],
            "tzo_reroot": [_P(TzTree), _vp, _vp, C.c_int],
            "tzo_synth_init_states": [_P(TzSynthGame), C.c_int, C.c_int, _vp, _vp, _vp],
            "tzo_synth_root": [_P(TzSynthGame), C.c_int, _vp, _vp, C.c_float, _vp, _vp],
            "tzo_synth_leaf": [_P(TzSynthGame), C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
            "tzo_synth_env_step": [_P(TzSynthGame), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp],
            "tzo_selfplay": [_P(TzTree), _P(TzSearchCfg), _P(TzSynthGame), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                             _vp, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int],
            "tzo_num_threads": [],
            "tzo_math": [C.c_int, _vp, C.c_float, _vp, C.c_int],
        }
        for name, args in sig.items():
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = args
    return _lib


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def make_cfg(selector=0, c=1.0, c1=1.25, c2=19652.0, epsilon=1e-8, discount=-1.0, weighted=False, q_temperature=1.0,
             fma_backup=False, q_transform=0) -> TzSearchCfg:
    inv_t = float(np.float32(1.0 / q_temperature)) if q_temperature > 0 else 0.0
    return TzSearchCfg(selector=selector, c=c, c1=c1, c2=c2, epsilon=epsilon, discount=discount, weighted=int(weighted),
                       inv_q_temperature=inv_t, fma_backup=int(fma_backup), q_transform=int(q_transform))


@dataclass
class HostTrees:
    """B trees in NumPy arrays with the device layout."""
    B: int
    N: int
    F: int
    emb_row_bytes: List[int]
    weighted: bool = False
    with_stats: bool = True
    next_free_idx: np.ndarray = field(init=False)
    parents: np.ndarray = field(init=False)
    edge_map: np.ndarray = field(init=False)
    n: np.ndarray = field(init=False)
    p: np.ndarray = field(init=False)
    q: np.ndarray = field(init=False)
    r: Optional[np.ndarray] = field(init=False)
    terminated: np.ndarray = field(init=False)
    emb: List[np.ndarray] = field(init=False)
    stats: Optional[np.ndarray] = field(init=False)

    def __post_init__(self):
        B, N, F = self.B, self.N, self.F
        self.next_free_idx = np.zeros((B,), np.int32)
        self.parents = np.full((B, N), -1, np.int32)
        self.edge_map = np.full((B, N, F), -1, np.int32)
        self.n = np.zeros((B, N), np.int32)
        self.p = np.zeros((B, N, F), np.float32)
        self.q = np.zeros((B, N), np.float32)
        self.r = np.zeros((B, N), np.float32) if self.weighted else None
        self.terminated = np.zeros((B, N), np.uint8)
        self.emb = [np.zeros((B, N, rb), np.uint8) for rb in self.emb_row_bytes]
        self.stats = np.zeros((B, 4), np.uint64) if self.with_stats else None

    def struct(self) -> TzTree:
        t = TzTree(B=self.B, N=self.N, F=self.F, n_emb=len(self.emb))
        t.next_free_idx = _ptr(self.next_free_idx)
        t.parents = _ptr(self.parents)
        t.edge_map = _ptr(self.edge_map)
        t.n = _ptr(self.n)
        t.p = _ptr(self.p)
        t.q = _ptr(self.q)
        t.r = _ptr(self.r)
        t.terminated = _ptr(self.terminated)
        for k, e in enumerate(self.emb):
            t.emb[k] = _ptr(e)
            t.emb_row_bytes[k] = self.emb_row_bytes[k]
        t.stats = _ptr(self.stats)
        return t

    def arrays(self) -> dict:
        d = {"next_free_idx": self.next_free_idx, "parents": self.parents, "edge_map": self.edge_map, "n": self.n,
             "p": self.p, "q": self.q, "terminated": self.terminated}
        if self.r is not None:
            d["r"] = self.r
        for k, e in enumerate(self.emb):
            d[f"emb{k}"] = e
        return d


@dataclass
class HostWork:
    B: int
    F: int
    emb_row_bytes: List[int]
    with_noise: bool = False

    def __post_init__(self):
        B, F = self.B, self.F
        self.parent = np.zeros((B,), np.int32)
        self.action = np.zeros((B,), np.int32)
        self.emb_parent = [np.zeros((B, rb), np.uint8) for rb in self.emb_row_bytes]
        self.policy = np.zeros((B, F), np.float32)
        self.value = np.zeros((B,), np.float32)
        self.terminated = np.zeros((B,), np.uint8)
        self.emb_new = [np.zeros((B, rb), np.uint8) for rb in self.emb_row_bytes]
        self.backprop_noise = np.zeros((B, F), np.float32) if self.with_noise else None

    def struct(self) -> TzWork:
        w = TzWork()
        w.parent = _ptr(self.parent)
        w.action = _ptr(self.action)
        w.policy = _ptr(self.policy)
        w.value = _ptr(self.value)
        w.terminated = _ptr(self.terminated)
        w.backprop_noise = _ptr(self.backprop_noise)
        w.path = None
        for k in range(len(self.emb_row_bytes)):
            w.emb_parent[k] = _ptr(self.emb_parent[k])
            w.emb_new[k] = _ptr(self.emb_new[k])
        return w


def _emb_ptrs(arrs: Sequence[np.ndarray]):
    arr = (_vp * max(len(arrs), 1))()
    for k, a in enumerate(arrs):
        arr[k] = _ptr(a)
    return arr


def _ok(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed: {rc}")


def set_root(t: HostTrees, root_policy, root_value, root_emb):
    _ok(lib().tzo_set_root(C.byref(t.struct()), _ptr(root_policy), _ptr(root_value), _emb_ptrs(root_emb)), "set_root")


def select(t: HostTrees, cfg: TzSearchCfg, w: HostWork):
    _ok(lib().tzo_select(C.byref(t.struct()), C.byref(cfg), C.byref(w.struct())), "select")


def expand_backprop(t: HostTrees, cfg: TzSearchCfg, w: HostWork):
    _ok(lib().tzo_expand_backprop(C.byref(t.struct()), C.byref(cfg), C.byref(w.struct())), "expand_backprop")


def root_action(t: HostTrees, temperature: float, noise=None, uniform01=None, want_action=True):
    B, F = t.B, t.F
    visits = np.zeros((B, F), np.int32)
    pw = np.zeros((B, F), np.float32)
    q0 = np.zeros((B,), np.float32)
    act = np.zeros((B,), np.int32) if want_action else None
    _ok(lib().tzo_root_action(C.byref(t.struct()), temperature, _ptr(noise), _ptr(uniform01), _ptr(visits), _ptr(pw),
                              _ptr(q0), _ptr(act)), "root_action")
    return act, pw, visits, q0


def reroot(t: HostTrees, action, reset_flag=None, persist_tree=True):
    _ok(lib().tzo_reroot(C.byref(t.struct()), _ptr(action), _ptr(reset_flag), int(persist_tree)), "reroot")


def make_game(F, payload_bytes, rho256, tau1024, max_depth, seed) -> TzSynthGame:
    return TzSynthGame(F=F, payload_bytes=payload_bytes, rho256=rho256, tau1024=tau1024, max_depth=max_depth, seed=seed)


def synth_init_states(g: TzSynthGame, B, env_offset, episode, core, payload):
    _ok(lib().tzo_synth_init_states(C.byref(g), B, env_offset, _ptr(episode), _ptr(core), _ptr(payload)), "synth_init")


def synth_root(g: TzSynthGame, B, core, dir_noise, dir_eps, root_policy, root_value):
    _ok(lib().tzo_synth_root(C.byref(g), B, _ptr(core), _ptr(dir_noise), dir_eps, _ptr(root_policy), _ptr(root_value)), "synth_root")


def synth_leaf(g: TzSynthGame, B, parent_core, action, policy, value, terminated, new_core, new_payload):
    _ok(lib().tzo_synth_leaf(C.byref(g), B, _ptr(parent_core), _ptr(action), _ptr(policy), _ptr(value), _ptr(terminated),
                             _ptr(new_core), _ptr(new_payload)), "synth_leaf")


def synth_env_step(g: TzSynthGame, B, env_offset, action, core, payload, episode, reset_flag):
    _ok(lib().tzo_synth_env_step(C.byref(g), B, env_offset, _ptr(action), _ptr(core), _ptr(payload), _ptr(episode),
                                 _ptr(reset_flag)), "synth_env_step")


def selfplay(t: HostTrees, cfg: TzSearchCfg, g: TzSynthGame, num_iterations, moves, temperature, persist_tree, env_offset,
             dir_noise, dir_eps, root_noise, uniform01, core, payload, episode, nthreads=0):
    """Tree-major whole self-play on the CPU (the timed CPU baseline).  Returns (actions [moves,B], pw [moves,B,F])."""
    actions = np.zeros((moves, t.B), np.int32)
    pw = np.zeros((moves, t.B, t.F), np.float32)
    if nthreads <= 0:
        nthreads = lib().tzo_num_threads()
    _ok(lib().tzo_selfplay(C.byref(t.struct()), C.byref(cfg), C.byref(g), num_iterations, moves, temperature,
                           int(persist_tree), env_offset, _ptr(dir_noise), dir_eps, _ptr(root_noise), _ptr(uniform01),
                           _ptr(core), _ptr(payload), _ptr(episode), _ptr(actions), _ptr(pw), nthreads), "selfplay")
    return actions, pw


def num_threads() -> int:
    return lib().tzo_num_threads()


def math(which: str, x: np.ndarray, y: float = 1.0) -> np.ndarray:
    """tz_expf / tz_logf / tz_powf of include/tz_math.h as compiled by gcc."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _ok(lib().tzo_math({"exp": 0, "log": 1, "pow": 2}[which], _ptr(x), y, _ptr(out), x.size), "math")
    return out
