"""Golden fixtures produced by the reference's own source (tests/golden/make_golden.py: /root/reference/core/... executed
on the NumPy emulation of the jax API, oracle/jaxshim) pin
  - both oracles (CPU, always), and
  - the CUDA path through the C-ABI (gpu-marked; needs neither /root/reference nor the oracle at run time).
Integers / bytes / indices bit-exact; floats compared with == (all three share the path's op order and exp/log)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as MG  # noqa: E402
from helpers import assert_trees_equal, check_invariants, run_c_stepwise, run_c_treemajor, run_numpy  # noqa: E402

NAMES = list(MG.CASES)


def _check(res, z, what, snapshots=True):
    assert np.array_equal(res.actions, z["actions"]), f"{what}: actions"
    assert np.array_equal(res.pw, z["pw"]), f"{what}: policy weights"
    if snapshots and res.snapshots is not None:
        for m, sn in enumerate(res.snapshots):
            assert_trees_equal(MG.split(z, f"search{m}"), sn, f"{what}: trees after the search of move {m}")
    assert_trees_equal(MG.split(z, "final"), res.arrays, f"{what}: final trees")


def test_fixture_files_exist_for_every_case():
    for name in NAMES:
        assert os.path.isfile(os.path.join(HERE, "golden", f"{name}.npz")), name


@pytest.mark.parametrize("name", NAMES)
def test_numpy_oracle_reproduces_reference_fixture(name):
    s, z = MG.load_case(name)
    _check(run_numpy(s, snapshots=True), z, name)


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_reproduces_reference_fixture(name):
    s, z = MG.load_case(name)
    res = run_c_stepwise(s, snapshots=True)
    _check(res, z, name)
    check_invariants(res.arrays)
    if s.bp_noise is None:
        _check(run_c_treemajor(s, nthreads=2), z, name + " (tree-major)")


def test_fixtures_regenerate_identically_when_the_reference_is_present():
    """Only in the build container: re-run the reference source on the shim and compare with the committed bytes."""
    from oracle import ref_via_shim as RV

    if not RV.available():
        pytest.skip("/root/reference is not mounted here")
    for name in ("ttt_T0_fulltree", "weighted_qT0", "short_episodes_persist"):
        out = MG.case_to_npz(name)
        old = dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))
        assert set(out) == set(old)
        for k in out:
            assert np.array_equal(out[k], old[k]), (name, k)


# ---- CUDA path vs the fixtures (no oracle involved) ----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_api_reproduces_reference_fixture(name):
    from helpers import run_cuda_api

    s, z = MG.load_case(name)
    res = run_cuda_api(s, fused=True, snapshots=True)
    _check(res, z, name)
    check_invariants(res.arrays)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if n != "weighted_qT0"])
@pytest.mark.parametrize("graph", [False, True])
def test_cuda_c_loop_reproduces_reference_fixture(name, graph):
    from helpers import run_cuda_selfplay

    s, z = MG.load_case(name)
    _check(run_cuda_selfplay(s, graph=graph), z, name, snapshots=False)
