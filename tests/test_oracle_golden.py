"""Pin the CPU oracle against the reference: golden k values from the reference's own tests and
phi / psi / alpha vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import ae_oracle as orc
from tests.kp import kp_problem, kp_psi0

# reference goldens: tests/test_short_kornreich_parsons.py:39-42, tests/long_kornreich_parsons.py:35-36
REF_GOLDEN_K = {(50, 4): 0.9358157950764868, (100, 16): 0.9885264464200719, (200, 196): 0.9900590623010896}


@pytest.mark.parametrize("cpm,A", [(50, 4), (100, 16), (200, 196)])
def test_reference_golden_k(cpm, A):
    p = kp_problem(9, cpm, A)
    r = orc.power_iteration(p["mu"], p["weight"], p["hx"], p["tables"], p["phi0"])
    assert r["status"] == 0
    assert abs(r["k_old"] - REF_GOLDEN_K[(cpm, A)]) < 1e-8      # the reference's own tolerance
    assert abs(r["k_old"] - REF_GOLDEN_K[(cpm, A)]) < 1e-13     # what the restatement actually achieves


def test_k_fixtures_bit_level(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "kp_k_*.npz")))
    assert len(files) >= 5
    for f in files:
        g = np.load(f)
        p = kp_problem(int(g["length"]), int(g["cells_per_mfp"]), int(g["num_angles"]))
        np.testing.assert_array_equal(p["phi0"], g["phi0"])             # same RNG draw as phi.py:2,10-11
        np.testing.assert_array_equal(p["total_xs"], g["total_xs"])     # same material map as slab.py:41-56
        np.testing.assert_array_equal(p["scatter_xs"], g["scatter_xs"])
        np.testing.assert_array_equal(p["nu_sigmaf"], g["nu_sigmaf"])
        r = orc.power_iteration(p["mu"], p["weight"], p["hx"], p["tables"], p["phi0"], want_psi=True)
        assert r["status"] == 0
        assert list(r["si_hist"]) == list(g["si_hist"])                 # same iteration counts
        assert abs(r["k_new"] - float(g["k_new"])) < 1e-14
        np.testing.assert_allclose(r["phi_new"], g["phi_new"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(r["phi_old"], g["phi_old"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(r["psi"][:, :, 0], g["psi_midpoint"], rtol=1e-13, atol=1e-300)


def test_sweep_fixtures_bit_exact(golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "kp_sweep_*.npz"))):
        g = np.load(f)
        mu, hx = g["mu"], float(g["hx"])
        for n in range(len(mu)):
            psi = orc.sweep(mu[n], hx, g["source"], g["total_xs"])
            np.testing.assert_array_equal(psi, g["psi"][:, n])          # bit-exact single sweep, group.py:41-60
        p = kp_problem(int(g["length"]), int(g["cells_per_mfp"]), int(g["num_angles"]))
        st, phi, it, psi = orc.source_iteration(mu, g["weight"], hx, p["tables"], g["source"], want_psi=True)
        assert st == 0
        np.testing.assert_allclose(phi, g["phi_si"], rtol=1e-14, atol=0)
        np.testing.assert_allclose(psi[:, :, 0], g["psi_si"], rtol=1e-14, atol=0)


def _sorted_c(a):
    a = np.asarray(a, dtype=complex)
    return a[np.lexsort((a.imag, -a.real))]


def test_dmd_fixtures(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "dmd_*.npz")))
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        if "psi" in g.files:
            psi = g["psi"]
        else:
            # answers-only fixture of a long march: its snapshots are the oracle's own (repaired) march, regenerated here
            # (K-P 50 cells/mfp S16: Z = 450); iteration counts and the last flux are part of the fixture
            cpm = int(g["Z"]) // 9
            p = kp_problem(9, cpm, int(g["A"]))
            res = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)), kp_psi0(p),
                                 float(g["dt"]), int(g["steps"]))
            assert res["status"] == 0
            np.testing.assert_array_equal(res["outer"], g["outer"])
            np.testing.assert_array_equal(res["inner"], g["inner"])
            np.testing.assert_array_equal(res["phi_last"], g["phi_last"])
            psi = res["psi_snap"]
        al = orc.dmd_alpha_from_psi(psi, int(g["steps"]), float(g["dt"]))
        ref = g["alphas"]
        assert len(al) == len(ref)                                      # same rank decision (time_unit.py:75-76)
        # same LAPACK, same memory layout as time_unit.py:68-72 -> identical to rounding
        np.testing.assert_allclose(_sorted_c(al), _sorted_c(ref), rtol=1e-12, atol=1e-13)


def test_dmd_march_matches_oracle_march(golden_dir):
    """The march fixture's snapshots came from the oracle's repaired time march: regenerate and compare."""
    g = np.load(os.path.join(golden_dir, "dmd_march_10_4_50.npz"))
    p = kp_problem(9, 10, 4)
    res = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)), kp_psi0(p), 0.1, 50)
    assert res["status"] == 0
    np.testing.assert_array_equal(res["psi_snap"], g["psi"])
    al = orc.dmd_alpha_from_psi(res["psi_snap"], 50, 0.1)
    dom = _sorted_c(al)[0]
    assert abs(dom - _sorted_c(g["alphas"])[0]) < 1e-10 * abs(dom)


@pytest.mark.parametrize("name", ["kp_march_repaired_9_5_2_6", "kp_march_repaired_9_10_4_4"])
def test_time_march_equals_repaired_reference(golden_dir, name):
    """The fixtures come from the reference's own sweep1D and time_dependent_source_iteration (unmodified) driven
    through the minimal repairs R2-R5 of its broken time loop (tests/golden/make_golden.py, mode `march`).  The
    oracle's march must reproduce them: same iteration counts, psi and phi snapshots to the last bit."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cpm, A, steps = int(g["cells_per_mfp"]), int(g["num_angles"]), int(g["steps"])
    p = kp_problem(9, cpm, A)
    np.testing.assert_array_equal(p["phi0"], g["phi0"])
    r = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)), kp_psi0(p), float(g["dt"]), steps)
    assert r["status"] == 0
    assert list(r["outer"]) == list(g["outer"]) and list(r["inner"]) == list(g["inner"])
    np.testing.assert_array_equal(r["psi_snap"].reshape(g["psi"].shape), g["psi"])
    np.testing.assert_array_equal(r["phi_snap"].reshape(g["phi"].shape), g["phi"])


def test_multigroup_source_terms_equal_reference_zone_formulas(golden_dir):
    """The only executable multigroup code of the reference, Zone.compute_fission_source / compute_scattering_source
    (zone.py:38-51), evaluated on seeded data (make_golden.py, mode `zone`).  The oracle's fission term must equal
    zone.py:38-42 exactly; zone.py:48-51 adds, per source group, the OUT-of-group scatter and the fission
    contribution, i.e. the oracle's full scatter term minus its in-group part plus its fission term."""
    g = np.load(os.path.join(golden_dir, "mg_zone_sources_7_40.npz"))
    G = int(g["G"])
    one = np.ones((3, G))
    tab = orc.Tables(g["mat"], one, g["scat"], g["chi"], g["nusf"], one)
    fis, sca = orc.mg_sources(tab, g["phi"])
    np.testing.assert_array_equal(fis, g["fission_source"])
    in_group = 0.5 * g["phi"] * np.array([np.diag(g["scat"][m]) for m in g["mat"]])
    ref = g["scatter_plus_fission_source"]
    assert np.max(np.abs((sca - in_group + fis) - ref) / np.abs(ref)) < 5e-15     # same terms, different summation order


def test_multigroup_reduces_to_one_group():
    """G=2 with two uncoupled identical groups must reproduce the one-group answer group by group."""
    p = kp_problem(9, 10, 4)
    t1 = p["tables"]
    M = t1.M
    scat = np.zeros((M, 2, 2))
    scat[:, 0, 0] = t1.scat[:, 0, 0]
    scat[:, 1, 1] = t1.scat[:, 0, 0]
    t2 = orc.Tables(t1.mat, np.repeat(t1.total, 2, 1), scat, np.repeat(t1.chi, 2, 1), np.repeat(t1.nusf, 2, 1),
                    np.repeat(t1.invspeed, 2, 1))
    src = np.random.RandomState(3).rand(p["Z"], 1) + 0.1
    st1, phi1, it1 = orc.source_iteration(p["mu"], p["weight"], p["hx"], t1, src)
    st2, phi2, it2 = orc.source_iteration(p["mu"], p["weight"], p["hx"], t2, np.repeat(src, 2, 1))
    assert st1 == st2 == 0 and it1 == it2
    np.testing.assert_array_equal(phi2[:, 0:1], phi1)
    np.testing.assert_array_equal(phi2[:, 1:2], phi1)


def test_not_converged_status():
    p = kp_problem(9, 5, 2)
    st, phi, it = orc.source_iteration(p["mu"], p["weight"], p["hx"], p["tables"], np.ones((p["Z"], 1)),
                                       max_iterations=5)
    assert st == 1 and it == 4          # `assert iteration < max_iterations` (group.py:72): 4 iterations run


def test_incident_boundary_flux_analytic():
    """OneDirection.boundary_condition (one_direction.py:12,24-31): with no source, the flux entering a uniform
    absorber is multiplied by a = (m - s/2)/(m + s/2) per cell of the diamond-difference recurrence (group.py:50), so
    the cell-centre fluxes are psi0 a^i (1 + a)/2 -- in both directions."""
    from oracle import ae_oracle as orc
    Z, hx, sig, psi0 = 40, 0.1, 1.3, 2.5
    lib = orc.lib()
    for mu in (0.7, -0.3):
        m = abs(mu) / hx
        a = (m - 0.5 * sig) / (m + 0.5 * sig)
        psi = np.empty(Z)
        src, s = np.zeros(Z), np.full(Z, sig)
        import ctypes
        lib.aeo_sweep_in.restype = None
        lib.aeo_sweep_in(ctypes.c_int64(Z), ctypes.c_double(mu), ctypes.c_double(1 / hx), orc._d(src), ctypes.c_int64(1), orc._d(s),
                         ctypes.c_int64(1), orc._d(psi), ctypes.c_int64(1), ctypes.c_double(psi0))
        want = psi0 * a ** np.arange(Z) * (1 + a) / 2
        if mu < 0:
            want = want[::-1]
        np.testing.assert_allclose(psi, want, rtol=1e-13)
    # vacuum is the zero-inflow case of the same routine
    psi_v = orc.sweep(0.7, hx, np.ones(Z), np.full(Z, sig))
    assert psi_v[0] == 0.5 * (0 + 1 / (0.7 / hx + 0.5 * sig))


# ------------------------------------------------------------------ reflective left boundary ----
@pytest.mark.parametrize("cpm,A", [(50, 4), (100, 16)])
def test_reflective_half_slab_reproduces_reference_golden_k(cpm, A):
    """The reference has no reflective boundary, but its K-P slab is symmetric about its centre (slab.py:41-56; phi0 =
    half + half[::-1], phi.py:10-11): the right half with a TRUE reflective left face (oracle: aeo_set_reflect_left, a
    direction enters with what its mirror left with) must reproduce the reference's golden k of the whole slab, with the
    same iteration history."""
    p = kp_problem(9, cpm, A)
    h = p["Z"] // 2
    t = p["tables"]
    full = orc.power_iteration(p["mu"], p["weight"], p["hx"], t, p["phi0"])
    half = orc.Tables(t.mat[h:], t.total, t.scat, t.chi, t.nusf, t.invspeed)
    orc.set_reflect_left(True)
    try:
        r = orc.power_iteration(p["mu"], p["weight"], p["hx"], half, p["phi0"][h:])
    finally:
        orc.set_reflect_left(False)
    assert r["status"] == 0
    assert abs(r["k_new"] - REF_GOLDEN_K[(cpm, A)]) < 1e-13
    assert list(r["si_hist"]) == list(full["si_hist"])
    np.testing.assert_allclose(r["phi_new"], full["phi_new"][h:], rtol=1e-13)


def test_reflective_march_equals_mirrored_march():
    """Time-dependent multigroup form of the same statement: the reflective half-slab march equals the right half of the
    march of the mirrored slab (psi(-x, mu) = psi(x, -mu) initially), iteration counts included."""
    rs = np.random.RandomState(11)
    Z, A, G, M, steps = 60, 6, 3, 2, 3
    mu, w = np.polynomial.legendre.leggauss(A)
    mat = (np.arange(Z) * M) // Z
    total = 0.5 + rs.rand(M, G)
    scat = 0.3 * rs.rand(M, G, G) * total[:, :, None] / G
    chi = np.tile(rs.dirichlet(np.ones(G)), (M, 1))
    nusf = 0.05 * rs.rand(M, G)
    inv = 0.5 + rs.rand(M, G)
    half = orc.Tables(mat, total, scat, chi, nusf, inv)
    mirror = lambda a: np.concatenate([a[::-1], a], axis=0)
    full = orc.Tables(mirror(mat), total, scat, chi, nusf, inv)
    psi0 = rs.rand(Z, A, G)
    S = rs.rand(Z, G)
    anti = np.array([int(np.argmin(np.abs(mu + m))) for m in mu])
    rf = orc.time_march(mu, w, 0.1, full, mirror(S), np.concatenate([psi0[::-1][:, anti, :], psi0], axis=0), 0.1, steps)
    orc.set_reflect_left(True)
    try:
        rh = orc.time_march(mu, w, 0.1, half, S, psi0, 0.1, steps)
    finally:
        orc.set_reflect_left(False)
    assert rf["status"] == 0 and rh["status"] == 0
    assert list(rf["outer"]) == list(rh["outer"]) and list(rf["inner"]) == list(rh["inner"])
    np.testing.assert_allclose(rh["psi_last"], rf["psi_last"][Z:], rtol=1e-12)
    np.testing.assert_allclose(rh["phi_last"], rf["phi_last"][Z:], rtol=1e-12)
