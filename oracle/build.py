"""Builds the C oracle (test infrastructure) into oracle/_build/libtz_oracle.so with gcc.

The reference is pure Python/JAX with no C sources, so there is no `oracle/_ref` to compile.
"""
from __future__ import annotations

import hashlib
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
OUT_DIR = HERE / "_build"
OUT = OUT_DIR / "libtz_oracle.so"
FLAGS = ["-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-shared", "-fPIC"]


def build(force: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    srcs = [HERE / "tz_oracle.c", *sorted((ROOT / "include").glob("*.h")), *sorted((ROOT / "standin" / "include").glob("*.h"))]
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for s in srcs:
        h.update(s.read_bytes())
    stamp = OUT_DIR / "libtz_oracle.sha256"
    if not force and OUT.exists() and stamp.exists() and stamp.read_text().strip() == h.hexdigest():
        return OUT
    # (standin/include/tz_synth.h: the synthetic game's inline definition, shared with the device stand-in)
    cmd = ["gcc", *FLAGS, f"-I{ROOT / 'include'}", f"-I{ROOT / 'standin' / 'include'}", str(HERE / "tz_oracle.c"), "-o", str(OUT), "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{res.stderr}")
    stamp.write_text(h.hexdigest() + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
