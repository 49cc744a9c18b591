#!/usr/bin/env python
"""Generates tests/golden/two_player_*.npz: two AlphaZero(MCTS) evaluators of different strength playing the synthetic
game against each other through the reference's UNMODIFIED core/common.py `two_player_game_step` (run on oracle/jaxshim;
see oracle/ref_via_shim.run_reference_two_player).  Exercises MCTS.step on the OPPONENT's tree (child usually absent:
tree.py:201-203 -> empty tree), get_value, the value-estimate discounting and the outcome bookkeeping (common.py:194-231).

    python tests/golden/make_golden_two_player.py [--check]
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

from oracle import ref_via_shim as R  # noqa: E402
from oracle import synth_numpy as SN  # noqa: E402

CASES = {
    # game kwargs, evaluator 1, evaluator 2, games, max_steps, seed
    "two_player_c4": (dict(F=7, payload_bytes=8, rho256=230, tau1024=90, max_depth=9, seed=41), dict(N=40, S=24), dict(N=24, S=10, c=1.5), 6, 10, 1),
    "two_player_ttt_T0": (dict(F=9, payload_bytes=0, rho256=154, tau1024=60, max_depth=7, seed=42), dict(N=16, S=20, temperature=0.0),
                          dict(N=32, S=12), 5, 8, 2),
}


def make(case):
    gkw, e1, e2, B, max_steps, seed = case
    g = SN.SynthGame(**gkw)
    rng = np.random.default_rng(seed)
    T = (max_steps // 2) * 2
    x = dict(p1_first=rng.random(B) < 0.5,
             dir_noise=rng.dirichlet([0.3] * g.F, size=(T, B)).astype(np.float32),
             root_noise=(rng.random((T, B, g.F), dtype=np.float32) * np.float32(1e-8)).astype(np.float32),
             uniform01=rng.random((T, B), dtype=np.float32))
    x["p1_first"][0], x["p1_first"][1] = True, False
    out = R.run_reference_two_player(g, e1, e2, x["p1_first"], max_steps, x["dir_noise"], x["root_noise"], x["uniform01"])
    return x, out


def main():
    check = "--check" in sys.argv
    for name, case in CASES.items():
        x, out = make(case)
        blob = {"in_" + k: v for k, v in x.items()}
        blob.update({"ref_" + k: v for k, v in out.items()})
        path = os.path.join(HERE, name + ".npz")
        if check:
            old = np.load(path)
            for k, v in blob.items():
                assert np.array_equal(old[k], v), f"{name}: {k} differs"
            print(f"{name}: matches")
        else:
            np.savez_compressed(path, **blob)
            print(f"{name}: wrote {os.path.getsize(path)} bytes; completed at end {out['completed'][-1].tolist()}, outcomes {out['outcomes'].tolist()}")


if __name__ == "__main__":
    main()
