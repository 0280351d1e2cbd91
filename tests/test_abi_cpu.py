"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the host
layout helpers agree with the library, and the host mirror of the reference classes keeps the
reference's RNG/geometry contract.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from alpha_eigen_b200 import _lib, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "alpha_eigen_b200.h")).read()
    declared = set(re.findall(r"\b(ae_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ae_abi_version() == 2


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_lib.AeCtrl) == 64
    assert C.sizeof(_lib.AeSlab) == 16 + 16 + 11 * 8 + 8
    assert C.sizeof(_lib.AeWork) == 14 * 8 + 16


@pytest.mark.parametrize("Z", [1, 7, 255, 256, 257, 450, 1800, 5000])
def test_layout_matches_library(Z):
    lib = _lib.load()
    idx = layout.storage_index(Z)
    assert [lib.ae_cell_offset(int(z)) for z in range(Z)] == list(idx)
    assert lib.ae_padded_cells(Z) == layout.padded_cells(Z)
    assert len(set(idx)) == Z and idx.max() < layout.padded_cells(Z)
    a = np.random.RandomState(0).rand(3, Z)
    np.testing.assert_array_equal(layout.from_storage(layout.to_storage(a, Z), Z), a)


def test_partial_count():
    lib = _lib.load()
    assert lib.ae_partial_count(4, 2) == 2
    assert lib.ae_partial_count(16, 8) == 2
    assert lib.ae_sweep_sync_bytes() > 0
    assert lib.ae_partial_count(196, 98) == 14
    assert lib.ae_partial_count(64, 32) == 4
    assert lib.ae_partial_count(3, 0) == 1


def test_host_mirror_matches_reference_fixture(golden_dir):
    """Classes of the drop-in package reproduce phi0 (RNG order), geometry and cross-sections."""
    from alpha_eigen_modules.group import OneGroup
    from alpha_eigen_modules.material import SimpleMaterial
    from alpha_eigen_modules.slab import SymmetricHeterogeneousSlab
    for cpm, A in [(50, 4), (100, 16)]:
        g0 = np.load(os.path.join(golden_dir, f"kp_k_9_{cpm}_{A}.npz"))
        np.random.seed(2021)        # fresh-process state of phi.py:2
        f, m, a = SimpleMaterial('fuel', 1), SimpleMaterial('moderator', 1), SimpleMaterial('absorber', 1)
        s = SymmetricHeterogeneousSlab(9, cpm, A, m, fuel_mat=f, abs_mat=a)
        g = OneGroup(slab=s)
        g.create_zones()
        np.testing.assert_array_equal(g.phi.old, g0["phi0"])
        for name in ("total_xs", "scatter_xs", "chi", "nu_sigmaf"):
            np.testing.assert_array_equal(getattr(g, name), g0[name])
        np.testing.assert_array_equal(s.mu, g0["mu"])
        assert s.hx == float(g0["hx"])
        assert g.zone[0].region_name == "fuel" and g.zone[len(g.zone) // 2].region_name == "absorber"
        # the per-zone draws advance the RNG exactly like Z Zone() constructions (zone.py:11)
        after = np.random.rand()
        np.random.seed(2021)
        np.random.rand(s.num_zones, 1)
        for _ in range(s.num_zones):
            np.random.rand(1)
        assert after == np.random.rand()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alpha_eigen_b200.inputs import kornreich_parsons_symmetric
    with pytest.raises(_lib.AlphaEigenError):
        kornreich_parsons_symmetric(9, 5, 2)


def test_dmd_window_matches_oracle():
    from alpha_eigen_b200.time_unit import dmd_window
    from oracle.ae_oracle import dmd_window as ow
    for n in (10, 20, 50, 200, 333):
        assert dmd_window(n) == ow(n)


def test_synthetic_generators_follow_the_survey_recipe():
    from alpha_eigen_b200.ensemble import synthetic_ensemble
    from alpha_eigen_b200.synthetic import region_index, synthetic_problem
    p = synthetic_problem(900, 8, 12)
    c = p["scat"].sum(axis=2) / p["total"]
    np.testing.assert_allclose(c[0], 0.8)
    np.testing.assert_allclose(c[2], 0.1)                        # absorber (material.py:17-19)
    assert np.all(p["nusf"][1:] == 0) and np.all(p["nusf"][0] > 0)
    np.testing.assert_allclose(p["chi"].sum(axis=1), 1.0)
    idx = region_index(900)                                       # fuel | mod | absorber | mod | fuel at 1,2,7,8 of 9
    assert list(np.flatnonzero(np.diff(idx))) == [99, 199, 699, 799]
    ens = synthetic_ensemble(5, 64, 4, 3)
    scale = [m["nusf"][0, 0] / ens[2]["nusf"][0, 0] for m in ens]
    np.testing.assert_allclose(scale, [0.9, 0.95, 1.0, 1.05, 1.1])
    assert np.allclose(ens[0]["phi0"], ens[0]["phi0"][::-1])     # symmetric initial guess (phi.py:11)
