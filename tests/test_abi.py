"""The C-ABI shared libraries load without a GPU and export every symbol include/*.h declares; struct layouts seen by
ctypes equal the C compiler's; argument validation answers before any launch.  No compute calls here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
INCLUDE = ROOT / "include"
SYNTH_INCLUDE = ROOT / "standin" / "include"


@pytest.fixture(scope="module")
def built():
    from turbozero_b200 import build

    build.build()
    from turbozero_b200 import _abi

    return _abi


@pytest.fixture(scope="module")
def standin():
    """The synthetic stand-in's library (test / bench infrastructure; its own package, standin/)."""
    from standin import build

    build.build()
    from standin import abi

    return abi


def declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = set()
    for m in re.finditer(r"^\s*(?:const\s+)?(?:int|uint64_t|char\s*\*|const char\s*\*)\s+\*?\s*(tz_[a-z0-9_]+)\s*\(", text, flags=re.M):
        names.add(m.group(1))
    return names


def test_headers_declare_what_ctypes_binds(built, standin):
    abi = declared_functions(INCLUDE / "tz_abi.h") | declared_functions(INCLUDE / "tz_replay.h")
    synth = declared_functions(SYNTH_INCLUDE / "tz_synth.h") - abi
    assert abi == set(built.TZ_SYMBOLS), abi ^ set(built.TZ_SYMBOLS)
    synth_exported = {n for n in synth if not n.startswith(("tz_synth_init_h", "tz_synth_step_h", "tz_synth_legal", "tz_synth_logit",
                                                            "tz_synth_terminal", "tz_synth_reward", "tz_synth_value",
                                                            "tz_synth_payload_word"))}
    assert synth_exported == set(standin.TZ_SYNTH_SYMBOLS), synth_exported ^ set(standin.TZ_SYNTH_SYMBOLS)


def test_libraries_export_every_declared_symbol(built, standin):
    lib = C.CDLL(str(built.LIB_DIR / "libtz_b200.so"))
    for name in declared_functions(INCLUDE / "tz_abi.h") | declared_functions(INCLUDE / "tz_replay.h"):
        assert hasattr(lib, name), f"libtz_b200.so does not export {name}"
    synth = C.CDLL(str(standin.LIB_DIR / "libtz_synth.so"))
    for name in standin.TZ_SYNTH_SYMBOLS:
        assert hasattr(synth, name), f"libtz_synth.so does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", str(built.LIB_DIR / "libtz_b200.so")], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(built.TZ_SYMBOLS) <= exported


def test_product_library_does_not_depend_on_oracle_or_synth(built):
    out = subprocess.run(["ldd", str(built.LIB_DIR / "libtz_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "tz_synth" not in out


def test_product_package_does_not_import_the_stand_in_or_the_oracle():
    """turbozero_b200 (the product) must not reach into standin/ (bench / test stand-in) or oracle/ (the checker)."""
    pkg = ROOT / "turbozero_b200"
    for f in pkg.glob("*.py"):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+(standin|oracle)\b", text, flags=re.M), f"{f.name} imports test infrastructure"
    assert not (pkg / "synthetic.py").exists()


def test_abi_version_and_strerror(built):
    lib = built.lib()
    assert lib.tz_abi_version() == built.TZ_ABI_VERSION == 7
    assert lib.tz_strerror(0) == b"ok"
    assert b"invalid" in lib.tz_strerror(-1)
    assert b"not supported" in lib.tz_strerror(-2)


def test_struct_layouts_match_the_c_compiler(built, standin, tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "tz_synth.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(TzTree), sizeof(TzSearchCfg), sizeof(TzWork), sizeof(TzSynthGame), sizeof(TzSynthCtx));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(TzTree, next_free_idx), offsetof(TzTree, child_stats), offsetof(TzTree, best), offsetof(TzTree, sel_state), offsetof(TzTree, emb), offsetof(TzTree, emb_row_bytes), offsetof(TzTree, stats));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(TzSearchCfg, discount), offsetof(TzSearchCfg, inv_q_temperature), offsetof(TzSearchCfg, fma_backup), offsetof(TzSearchCfg, programmatic), offsetof(TzSearchCfg, q_transform), offsetof(TzSearchCfg, sim_warps));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(TzWork, emb_parent), offsetof(TzWork, policy), offsetof(TzWork, emb_new), offsetof(TzWork, path), offsetof(TzWork, path_spill), offsetof(TzWork, path_spill_cap), offsetof(TzWork, timeline_slots), offsetof(TzWork, timeline));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", f"-I{INCLUDE}", f"-I{SYNTH_INCLUDE}", str(src), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    T, S, W = built.TzTree, built.TzSearchCfg, built.TzWork
    assert lines[0].split() == [str(C.sizeof(x)) for x in (T, S, W, standin.TzSynthGame, standin.TzSynthCtx)]
    assert lines[1].split() == [str(getattr(T, f).offset) for f in ("next_free_idx", "child_stats", "best", "sel_state", "emb", "emb_row_bytes", "stats")]
    assert lines[2].split() == [str(getattr(S, f).offset) for f in ("discount", "inv_q_temperature", "fma_backup", "programmatic", "q_transform", "sim_warps")]
    assert lines[3].split() == [str(getattr(W, f).offset) for f in ("emb_parent", "policy", "emb_new", "path", "path_spill", "path_spill_cap", "timeline_slots", "timeline")]


def test_argument_validation_needs_no_gpu(built):
    lib = built.lib()
    t = built.TzTree(B=0, N=4, F=3, n_emb=0)
    assert lib.tz_tree_init(C.byref(t), None) == -1  # TZ_EINVAL: B <= 0
    t = built.TzTree(B=1, N=4, F=3, n_emb=0)  # null array pointers
    cfg = built.TzSearchCfg(selector=0, c=1.0, c1=0.0, c2=1.0, epsilon=1e-8, discount=-1.0)
    w = built.TzWork()
    assert lib.tz_select(C.byref(t), C.byref(cfg), C.byref(w), None) == -1
    assert lib.tz_reroot(C.byref(t), None, None, 1, None) == -1
    assert lib.tz_search(C.byref(t), C.byref(cfg), C.byref(w), 4, None, None, None) == -1  # no leaf callback
    assert lib.tz_launch_count() == 0  # nothing was launched by any of the above


def test_replay_argument_validation_needs_no_gpu(built):
    """include/tz_replay.h entry points reject malformed descriptors before touching the device."""
    lib = built.lib()
    r = built.TzReplay(B=0, capacity=4, n_leaves=1, reward_leaf=0, reward_dim=2)
    assert lib.tz_replay_init(C.byref(r), None) == -1  # B <= 0
    r = built.TzReplay(B=2, capacity=4, n_leaves=1, reward_leaf=0, reward_dim=2)  # null state pointers
    assert lib.tz_replay_collect(C.byref(r), 0, None, None, None, None, None) == -1
    assert lib.tz_replay_count_valid(C.byref(r), None, None) == -1
    assert lib.tz_replay_gather(C.byref(r), None, 3, None, None) == -1
    assert C.sizeof(built.TzReplay) == 6 * 4 + 4 * 8 + built.TZ_MAX_EMB * 16


def test_replay_struct_layout_matches_header(built, tmp_path):
    src = tmp_path / "layout_replay.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "tz_replay.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(TzReplay), offsetof(TzReplay, reward_dim), offsetof(TzReplay, next_idx),
         offsetof(TzReplay, has_reward), offsetof(TzReplay, leaf), offsetof(TzReplay, leaf_row_bytes));
  return 0;
}''')
    exe = tmp_path / "layout_replay"
    subprocess.run(["gcc", f"-I{INCLUDE}", str(src), "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    R = built.TzReplay
    assert got == [str(C.sizeof(R))] + [str(getattr(R, f).offset) for f in ("reward_dim", "next_idx", "has_reward", "leaf", "leaf_row_bytes")]


def test_missing_library_is_a_loud_error(built, monkeypatch, tmp_path):
    monkeypatch.setattr(built, "LIB_DIR", tmp_path)
    with pytest.raises(built.TzError, match="no CPU or PyTorch fallback"):
        built._load("libtz_b200.so", built.TZ_SYMBOLS)


def test_jax_ffi_adapter_is_gated_and_its_source_is_well_formed():
    """turbozero_b200/ffi_jax.py + csrc/tz_jax_ffi.cc (the jax.ffi registration the north star names) cannot be built here
    (no jax): the module must say so on import, and the C++ must at least be well-formed against a mock of the XLA FFI
    surface it uses (tests/mock_xla -- this checks OUR code for errors, not XLA's API)."""
    import importlib
    import sys

    real_jax = "jax" in sys.modules and not getattr(sys.modules["jax"], "__shim__", False)
    if not real_jax:
        sys.modules.pop("turbozero_b200.ffi_jax", None)
        with pytest.raises(ImportError, match="needs"):
            importlib.import_module("turbozero_b200.ffi_jax")
    root = INCLUDE.parent
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{INCLUDE}", f"-I{root / 'tests' / 'mock_xla'}",
                          "-I/usr/local/cuda/include", str(root / "turbozero_b200" / "csrc" / "tz_jax_ffi.cc")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
