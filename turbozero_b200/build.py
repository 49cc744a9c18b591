"""In-tree build of the CUDA libraries (nvcc cross-compiles sm_100a without a GPU).

    python -m turbozero_b200.build            # libtz_b200.so into turbozero_b200/lib/

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
]

# one translation unit per kernel family / register-chunk count: they compile in parallel (the per-simulation kernels are
# templates with ~100 instantiations; one TU took 3 minutes)
TARGETS = {
    "libtz_b200.so": ["csrc/tz_kernels.cu", "csrc/tz_replay.cu", "csrc/tz_reroot.cu", "csrc/tz_sim_nc1.cu", "csrc/tz_sim_nc2.cu",
                      "csrc/tz_sim_nc3.cu", "csrc/tz_sim_nc4.cu", "csrc/tz_sim_nc8.cu", "csrc/tz_sim_nc16.cu",
                      "csrc/tz_wide_plain.cu", "csrc/tz_wide_weighted.cu"],
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


def _compile(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, res


def build(force: bool = False, verbose: bool = False, extra_flags=(), suffix: str = "") -> None:
    """Compiles every translation unit of every target to an object file (in parallel) and links the shared libraries.
    `extra_flags` / `suffix` serve the diagnostic build (scripts/build_prof.sh: -DTZ_PROFILE -rdc=true -> lib*_prof.so)."""
    from concurrent.futures import ThreadPoolExecutor

    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / ("obj" + suffix)
    headers = sorted(INCLUDE.glob("*.h")) + sorted((PKG / "csrc").glob("*.cuh"))
    for out_name, srcs in TARGETS.items():
        out = LIB_DIR / out_name.replace(".so", suffix + ".so")
        src_paths = [PKG / s for s in srcs]
        stamp = LIB_DIR / (out.name + ".sha256")
        want = (_digest(src_paths + headers) + " " + " ".join(extra_flags)).strip()  # (compared with the stamp's stripped text)
        if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == want:
            continue
        obj_dir.mkdir(exist_ok=True)
        jobs = []
        for sp in src_paths:
            obj = obj_dir / (sp.stem + ".o")
            cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, f"-I{INCLUDE}", "-c", str(sp), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((cmd, obj))
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            results = list(pool.map(_compile, [j[0] for j in jobs]))
        for cmd, res in results:
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {out_name}:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
                print(res.stderr, file=sys.stderr)
        link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", *[f for f in extra_flags if f.startswith("-rdc")],
                "-Xcompiler", "-fPIC", "-shared", *[str(j[1]) for j in jobs], "-o", str(out)]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed for {out_name}:\n{' '.join(link)}\n{res.stdout}\n{res.stderr}")
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
    if "--exp" in sys.argv:  # experiment builds: python -m turbozero_b200.build --exp _notl -DTZ_NO_TIMELINE  -> lib*_notl.so
        i = sys.argv.index("--exp")
        build(force=True, extra_flags=tuple(sys.argv[i + 2:]), suffix=sys.argv[i + 1])
    elif "--prof" in sys.argv:  # diagnostic build with in-kernel phase clocks: lib*_prof.so (scripts/phase_*.py, scripts/timeline.py)
        build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=("-DTZ_PROFILE", "-rdc=true"), suffix="_prof")
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
