"""In-tree build of the CUDA libraries (nvcc cross-compiles sm_100a without a GPU).

    python -m turbozero_b200.build            # libtz_b200.so + libtz_synth.so into turbozero_b200/lib/

The built .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
INCLUDE = ROOT / "include"
LIB_DIR = PKG / "lib"

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # every float op individually rounded: reference op order, bit-exact vs the oracle
    "-Xcompiler", "-fPIC",
    "-shared",
]

TARGETS = {
    "libtz_b200.so": ["csrc/tz_kernels.cu", "csrc/tz_replay.cu"],
    "libtz_synth.so": ["csrc/tz_synth.cu"],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the sm_100a kernels cannot be built")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(Path(p).read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> None:
    LIB_DIR.mkdir(exist_ok=True)
    headers = sorted(INCLUDE.glob("*.h"))
    for out_name, srcs in TARGETS.items():
        out = LIB_DIR / out_name
        src_paths = [PKG / s for s in srcs]
        stamp = LIB_DIR / (out_name + ".sha256")
        want = _digest(src_paths + headers)
        if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == want:
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE}", *map(str, src_paths), "-o", str(out)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {out_name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        stamp.write_text(want + "\n")


def build_jax_ffi() -> Path:
    """csrc/tz_jax_ffi.cc -> lib/libtz_jax_ffi.so (the XLA FFI handlers of turbozero_b200/ffi_jax.py).  Needs jax for the XLA
    FFI headers (`jax.ffi.include_dir()`); NOT buildable in the image this repo was developed in (no jax, no network)."""
    try:
        try:
            from jax import ffi as jffi
        except ImportError:
            from jax.extend import ffi as jffi
    except ImportError as e:
        raise RuntimeError("build_jax_ffi() needs jax >= 0.4.35 for the XLA FFI headers") from e
    build()
    out = LIB_DIR / "libtz_jax_ffi.so"
    cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", f"-I{INCLUDE}", f"-I{jffi.include_dir()}",
           str(PKG / "csrc" / "tz_jax_ffi.cc"), f"-L{LIB_DIR}", "-l:libtz_b200.so", f"-Xlinker=-rpath={LIB_DIR}", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for libtz_jax_ffi.so:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
