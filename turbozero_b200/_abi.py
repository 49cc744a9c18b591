"""ctypes binding of include/tz_abi.h (libtz_b200.so) -- the only way this package reaches the kernels.

There is NO fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

TZ_MAX_EMB = 24
TZ_PATH_CAP = 32
TZ_PATH_STRIDE = 2 * TZ_PATH_CAP + 2
TZ_SEL_STATE_WORDS = 8
TZ_ABI_VERSION = 7
TZ_SEL_PUCT = 0
TZ_SEL_MUZERO_PUCT = 1
TZ_QT_NORMALIZE = 0
TZ_QT_IDENTITY = 1

LIB_DIR = Path(__file__).resolve().parent / "lib"


class TzTree(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("F", C.c_int32), ("n_emb", C.c_int32),
        ("next_free_idx", C.c_void_p), ("parents", C.c_void_p), ("edge_map", C.c_void_p),
        ("n", C.c_void_p), ("p", C.c_void_p), ("q", C.c_void_p), ("r", C.c_void_p),
        ("terminated", C.c_void_p), ("child_stats", C.c_void_p), ("best", C.c_void_p), ("sel_state", C.c_void_p),
        ("emb", C.c_void_p * TZ_MAX_EMB),
        ("emb_row_bytes", C.c_int64 * TZ_MAX_EMB),
        ("stats", C.c_void_p),
    ]


class TzSearchCfg(C.Structure):
    _fields_ = [
        ("selector", C.c_int32), ("c", C.c_float), ("c1", C.c_float), ("c2", C.c_float),
        ("epsilon", C.c_float), ("discount", C.c_float), ("weighted", C.c_int32),
        ("inv_q_temperature", C.c_float), ("fma_backup", C.c_int32), ("programmatic", C.c_int32),
        ("q_transform", C.c_int32), ("sim_warps", C.c_int32),
    ]


class TzWork(C.Structure):
    _fields_ = [
        ("parent", C.c_void_p), ("action", C.c_void_p),
        ("emb_parent", C.c_void_p * TZ_MAX_EMB),
        ("policy", C.c_void_p), ("value", C.c_void_p), ("terminated", C.c_void_p),
        ("emb_new", C.c_void_p * TZ_MAX_EMB),
        ("backprop_noise", C.c_void_p), ("path", C.c_void_p),
        ("path_spill", C.c_void_p), ("path_spill_cap", C.c_int32), ("timeline_slots", C.c_int32),
        ("timeline", C.c_void_p),
    ]


class TzReplay(C.Structure):
    """include/tz_replay.h"""
    _fields_ = [
        ("B", C.c_int32), ("capacity", C.c_int32), ("n_leaves", C.c_int32), ("reward_leaf", C.c_int32),
        ("reward_dim", C.c_int32), ("pad", C.c_int32),
        ("next_idx", C.c_void_p), ("episode_start_idx", C.c_void_p), ("populated", C.c_void_p), ("has_reward", C.c_void_p),
        ("leaf", C.c_void_p * TZ_MAX_EMB),
        ("leaf_row_bytes", C.c_int64 * TZ_MAX_EMB),
    ]


LEAF_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(TzWork), C.c_void_p)

_P = C.POINTER
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/tz_abi.h declares
TZ_SYMBOLS = {
    "tz_abi_version": (C.c_int, []),
    "tz_strerror": (C.c_char_p, [C.c_int]),
    "tz_launch_count": (C.c_uint64, []),
    "tz_launch_seq": (C.c_uint64, []),
    "tz_tree_init": (C.c_int, [_P(TzTree), _vp]),
    "tz_rebuild_child_stats": (C.c_int, [_P(TzTree), _vp]),
    "tz_set_root": (C.c_int, [_P(TzTree), _vp, _vp, _P(_vp), _vp]),
    "tz_select": (C.c_int, [_P(TzTree), _P(TzSearchCfg), _P(TzWork), _vp]),
    "tz_expand_backprop": (C.c_int, [_P(TzTree), _P(TzSearchCfg), _P(TzWork), _vp]),
    "tz_expand_backprop_select": (C.c_int, [_P(TzTree), _P(TzSearchCfg), _P(TzWork), _vp]),
    "tz_root_action": (C.c_int, [_P(TzTree), C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tz_reroot": (C.c_int, [_P(TzTree), _vp, _vp, C.c_int, _vp]),
    "tz_selftest_div": (C.c_int, [C.c_uint64, C.c_uint32, _vp, _vp]),
    "tz_selftest_best": (C.c_int, [_P(TzTree), _P(TzSearchCfg), _vp, _vp]),
    "tz_search": (C.c_int, [_P(TzTree), _P(TzSearchCfg), _P(TzWork), C.c_int, _vp, _vp, _vp]),
    # include/tz_replay.h
    "tz_replay_init": (C.c_int, [_P(TzReplay), _vp]),
    "tz_replay_collect": (C.c_int, [_P(TzReplay), C.c_int, _P(_vp), _vp, _vp, _vp, _vp]),
    "tz_replay_count_valid": (C.c_int, [_P(TzReplay), _vp, _vp]),
    "tz_replay_sample_scores": (C.c_int, [_P(TzReplay), _vp, _vp, _vp, _vp]),
    "tz_replay_gather": (C.c_int, [_P(TzReplay), _vp, C.c_int, _P(_vp), _vp]),
}

class TzError(RuntimeError):
    pass


def _load(name: str, symbols, lib_dir: Path = None) -> C.CDLL:
    path = (lib_dir or LIB_DIR) / name
    if not path.exists():
        raise TzError(
            f"{path} is missing: build it with `python -m turbozero_b200.build` (needs nvcc). "
            "turbozero_b200 has no CPU or PyTorch fallback for the search kernels.")
    lib = C.CDLL(str(path))
    for sym, (res, args) in symbols.items():
        fn = getattr(lib, sym)  # AttributeError here == ABI mismatch, which must be loud
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    """libtz_b200.so (loaded once)."""
    global _lib
    if _lib is None:
        # TZ_B200_LIB: file name of an alternative build in lib/ (diagnostic / experimental builds of the same ABI)
        _lib = _load(os.environ.get("TZ_B200_LIB", "libtz_b200.so"), TZ_SYMBOLS)
        if _lib.tz_abi_version() != TZ_ABI_VERSION:
            raise TzError("libtz_b200.so ABI version mismatch")
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().tz_strerror(rc)
        raise TzError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")
