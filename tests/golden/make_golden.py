#!/usr/bin/env python
"""Generates tests/golden/*.npz by executing the reference's UNMODIFIED source (/root/reference/core/...) on the NumPy
emulation of the jax API in oracle/jaxshim (jax is not installable in this image; oracle/jaxshim/README.md says exactly
what that does and does not pin).  Each fixture stores the schedule's inputs (so no RNG stream has to be reproducible),
and the reference's outputs: per-move actions and policy weights, the trees after every search (before re-rooting) and
the final trees.

    python tests/golden/make_golden.py            # regenerate everything (needs /root/reference)
    python tests/golden/make_golden.py --check    # regenerate in memory and compare with the committed files
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import Schedule  # noqa: E402
from oracle import synth_numpy as SN  # noqa: E402


def G(**kw):
    return SN.SynthGame(**kw)


# name -> Schedule kwargs.  Small on purpose (committed fixtures); shapes follow BASELINE.json's configs.
CASES = {
    # configs[0]: tic_tac_toe AlphaZero MCTS (F=9); argmax root action, tiny trees that fill up (Q7) and re-root every move
    "ttt_T0_fulltree": dict(game=G(F=9, payload_bytes=7, rho256=154, tau1024=40, max_depth=9, seed=1), B=4, N=12, S=30, moves=6,
                            temperature=0.0),
    "ttt_cfg1_shape": dict(game=SN.make_game("tic_tac_toe", 1001), B=3, N=128, S=64, moves=4, temperature=1.0),
    # configs[1]: connect_four shape, subtree persistence on
    "c4_persist": dict(game=G(F=7, payload_bytes=32, rho256=230, tau1024=12, max_depth=42, seed=2), B=4, N=64, S=40, moves=5,
                       temperature=1.0),
    "c4_cfg2_shape": dict(game=SN.make_game("connect_four", 2000), B=2, N=256, S=128, moves=3, temperature=1.0),
    # configs[2]: othello shape, WeightedMCTS backup (weighted_mcts.py) at three q-temperatures incl. the argmax branch
    "othello_weighted": dict(game=G(F=65, payload_bytes=16, rho256=38, tau1024=6, max_depth=60, seed=3), B=2, N=50, S=30, moves=3,
                             temperature=1.0, weighted=True),
    "weighted_qT05": dict(game=G(F=33, payload_bytes=5, rho256=60, tau1024=6, max_depth=60, seed=3), B=2, N=50, S=30, moves=3,
                          temperature=0.5, weighted=True, q_temperature=0.5),
    "weighted_qT0": dict(game=G(F=33, payload_bytes=5, rho256=100, tau1024=6, max_depth=60, seed=4), B=2, N=50, S=30, moves=3,
                         temperature=1.0, weighted=True, q_temperature=0.0),
    # configs[3]: go_9x9 shape (82-way), deeper trees
    "go_F82": dict(game=G(F=82, payload_bytes=48, rho256=205, tau1024=2, max_depth=120, seed=5), B=2, N=80, S=60, moves=2,
                   temperature=1.0),
    # configs[4]: 2048 shape: single player, discount +1, plain MCTS root (no Dirichlet noise, mcts.py:123-142)
    "g2048_pos_discount": dict(game=G(F=4, payload_bytes=16, rho256=218, tau1024=4, max_depth=200, seed=6), B=4, N=40, S=30,
                               moves=5, temperature=1.0, discount=1.0, dirichlet=False),
    # MCTS.step with persist_tree=False (mcts.py:413-414) and many episode ends (common.py:89-100)
    "no_persist_short_episodes": dict(game=G(F=7, payload_bytes=3, rho256=230, tau1024=100, max_depth=6, seed=7), B=3, N=64,
                                      S=40, moves=6, temperature=1.0, persist_tree=False),
    "short_episodes_persist": dict(game=G(F=5, payload_bytes=4, rho256=200, tau1024=150, max_depth=5, seed=13), B=4, N=48, S=32,
                                   moves=8, temperature=1.0),
    # deep narrow trees, fractional discount, c != 1
    "deep_discount09": dict(game=G(F=5, payload_bytes=3, rho256=230, tau1024=0, max_depth=100, seed=8), B=2, N=200, S=150,
                            moves=2, temperature=1.0, discount=0.9, c=2.5),
    # MuZeroPUCTSelector (action_selection.py:119-177), run through the reference's unmodified class with a five-argument
    # adapter around its own normalize_q_values as `q_transform` (the arity bug of :169; see oracle/ref_via_shim.py)
    "muzero_puct_arityfix": dict(game=G(F=7, payload_bytes=8, rho256=230, tau1024=12, max_depth=42, seed=21), B=3, N=64, S=48,
                                 moves=4, temperature=1.0, selector=1, c1=1.25, c2=19652.0),
    "muzero_puct_arityfix_wide_small_c2": dict(game=G(F=40, payload_bytes=4, rho256=120, tau1024=6, max_depth=60, seed=22), B=2,
                                               N=72, S=60, moves=3, temperature=0.0, selector=1, c1=0.7, c2=11.0),
    # PUCTSelector with a different registered q_transform: the identity (include/tz_abi.h TZ_QT_IDENTITY)
    "puct_identity_qtransform": dict(game=G(F=7, payload_bytes=8, rho256=230, tau1024=12, max_depth=42, seed=23), B=3, N=64, S=48,
                                     moves=4, temperature=1.0, q_transform=1),
    "puct_identity_qtransform_wide": dict(game=G(F=36, payload_bytes=0, rho256=140, tau1024=6, max_depth=60, seed=24), B=2, N=60,
                                          S=50, moves=3, temperature=1.0, q_transform=1, discount=0.9),
    "very_deep_F2": dict(game=G(F=2, payload_bytes=4, rho256=0, tau1024=0, max_depth=1000, seed=12), B=2, N=120, S=100, moves=2,
                         temperature=1.0, discount=0.97),
}

INPUT_KEYS = ("dir_noise", "root_noise", "uniform01", "bp_noise")


def case_to_npz(name: str) -> dict:
    from oracle import ref_via_shim as RV

    kw = CASES[name]
    s = Schedule(**kw)
    arrays, actions, pw, snaps = RV.run_reference(s, snapshots=True)
    out = {"actions": actions, "pw": pw}
    for k, v in arrays.items():
        out[f"final/{k}"] = v
    for m, sn in enumerate(snaps):
        for k, v in sn.items():
            out[f"search{m}/{k}"] = v
    for k in INPUT_KEYS:
        v = getattr(s, k)
        if v is not None:
            out[f"in/{k}"] = v
    return out


def load_case(name: str):
    """(Schedule with the stored inputs, fixture dict) for tests."""
    kw = CASES[name]
    s = Schedule(**kw)
    z = dict(np.load(os.path.join(HERE, f"{name}.npz")))
    for k in INPUT_KEYS:
        setattr(s, k, z.get(f"in/{k}"))
    return s, z


def split(z: dict, prefix: str) -> dict:
    return {k[len(prefix) + 1:]: v for k, v in z.items() if k.startswith(prefix + "/")}


def main():
    check = "--check" in sys.argv
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    bad = 0
    for name in names:
        out = case_to_npz(name)
        path = os.path.join(HERE, f"{name}.npz")
        if check:
            old = dict(np.load(path))
            same = set(old) == set(out) and all(np.array_equal(old[k], out[k]) for k in out)
            print(f"{name}: {'ok' if same else 'DIFFERS'}")
            bad += not same
        else:
            np.savez_compressed(path, **out)
            print(f"{name}: wrote {os.path.getsize(path)} bytes")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
