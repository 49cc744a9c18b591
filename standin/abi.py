"""ctypes binding of standin/include/tz_synth.h (libtz_synth.so): the synthetic game stand-in.  Bench / test infrastructure,
not part of the product package (turbozero_b200 never imports this)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from turbozero_b200._abi import TzWork, _load

LIB_DIR = Path(__file__).resolve().parent / "lib"

_P = C.POINTER
_vp = C.c_void_p


class TzSynthGame(C.Structure):
    _fields_ = [
        ("F", C.c_int32), ("payload_bytes", C.c_int32), ("rho256", C.c_int32), ("tau1024", C.c_int32),
        ("max_depth", C.c_int32), ("seed", C.c_uint32),
    ]


class TzSynthCtx(C.Structure):
    _fields_ = [("game", TzSynthGame), ("B", C.c_int32)]



TZ_SYNTH_SYMBOLS = {
    "tz_synth_launch_count": (C.c_uint64, []),
    "tz_synth_init_states": (C.c_int, [_P(TzSynthGame), C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "tz_synth_root": (C.c_int, [_P(TzSynthGame), C.c_int, _vp, _vp, C.c_float, _vp, _vp, _vp]),
    "tz_synth_leaf": (C.c_int, [_P(TzSynthGame), C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tz_synth_env_step": (C.c_int, [_P(TzSynthGame), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tz_synth_leaf_cb": (C.c_int, [_vp, C.c_int, _P(TzWork), _vp]),
    "tz_synth_timed_begin": (C.c_int, [C.c_int]),
    "tz_synth_leaf_cb_timed": (C.c_int, [_vp, C.c_int, _P(TzWork), _vp]),
    "tz_synth_timed_collect": (C.c_int, [_vp, _vp]),
    "tz_synth_set_programmatic": (C.c_int, [C.c_int]),
    "tz_synth_set_timeline": (C.c_int, [_vp, C.c_int]),
    "tz_synth_leaf_seq": (C.c_uint64, []),
}


_synth = None


def synth_lib() -> C.CDLL:
    """libtz_synth.so (loaded once)."""
    global _synth
    if _synth is None:
        _synth = _load("libtz_synth.so", TZ_SYNTH_SYMBOLS, LIB_DIR)
    return _synth
