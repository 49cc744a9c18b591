"""In-tree build of libtz_synth.so (the synthetic stand-in; nvcc cross-compiles sm_100a without a GPU).

    python -m standin.build [--prof]
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

from turbozero_b200.build import INCLUDE, NVCC_FLAGS, _digest, _nvcc

PKG = Path(__file__).resolve().parent
LIB_DIR = PKG / "lib"
SRC = PKG / "csrc" / "tz_synth.cu"
HDR = PKG / "include"


def build(force: bool = False, extra_flags=(), suffix: str = "") -> None:
    LIB_DIR.mkdir(exist_ok=True)
    out = LIB_DIR / f"libtz_synth{suffix}.so"
    stamp = LIB_DIR / (out.name + ".sha256")
    want = (_digest([SRC] + sorted(HDR.glob("*.h")) + sorted(INCLUDE.glob("*.h"))) + " " + " ".join(extra_flags)).strip()
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return
    cmd = [_nvcc(), *NVCC_FLAGS, *[f for f in extra_flags if not f.startswith("-rdc")], f"-I{INCLUDE}", f"-I{HDR}", "-shared", str(SRC), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {out.name}:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    stamp.write_text(want + "\n")


if __name__ == "__main__":
    if "--prof" in sys.argv:
        build(force="--force" in sys.argv, extra_flags=("-DTZ_PROFILE",), suffix="_prof")
    else:
        build(force="--force" in sys.argv)
