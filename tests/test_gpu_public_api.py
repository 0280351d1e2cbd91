"""GPU tests of the drop-in surface: the reference's classes, CLI and the boundary helpers of the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _kp(cpm, A):
    from alpha_eigen_modules.group import OneGroup
    from alpha_eigen_modules.material import SimpleMaterial
    from alpha_eigen_modules.slab import SymmetricHeterogeneousSlab
    np.random.seed(2021)
    f, m, a = SimpleMaterial('fuel', 1), SimpleMaterial('moderator', 1), SimpleMaterial('absorber', 1)
    s = SymmetricHeterogeneousSlab(9, cpm, A, m, fuel_mat=f, abs_mat=a)
    g = OneGroup(slab=s)
    g.create_zones()
    return s, g


def test_reference_test_body_runs_unchanged():
    """The body of the reference's own test (tests/test_short_kornreich_parsons.py:17-42), through the
    alias package, both configurations in one process like the reference does."""
    from alpha_eigen_modules.group import OneGroup
    from alpha_eigen_modules.material import SimpleMaterial
    from alpha_eigen_modules.slab import SymmetricHeterogeneousSlab

    def calculate_k_eigenvalue(length, cells_per_mfp, num_angles):
        fuel = SimpleMaterial('fuel', num_groups=1)
        moderator = SimpleMaterial('moderator', num_groups=1)
        absorber = SimpleMaterial('absorber', num_groups=1)
        slab = SymmetricHeterogeneousSlab(length, cells_per_mfp, num_angles, moderator, fuel_mat=fuel, abs_mat=absorber)
        one_group = OneGroup(slab=slab)
        one_group.create_zones()
        return one_group.perform_power_iteration(quiet=True)

    np.random.seed(2021)
    k = calculate_k_eigenvalue(length=9, cells_per_mfp=50, num_angles=4)
    assert abs(k.old - 0.9358157950764868) < 1e-8
    k = calculate_k_eigenvalue(length=9, cells_per_mfp=100, num_angles=16)     # RNG state now differs from a fresh process
    assert abs(k.old - 0.9885264464200719) < 1e-8


def test_cli_matches_reference_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--problem", "kornreich-parsons-symmetric",
                          "--length", "9", "--cells_per_mfp", "10", "--num_angles", "4"], capture_output=True, text=True,
                         cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Initiating k-eigenvalue solver" in out.stdout and "Power Iteration 1 : k =" in out.stdout
    assert abs(float(out.stdout.strip().splitlines()[-1]) - 0.934364027690702) < 1e-9      # oracle value for this mesh
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--problem", "kornreich-parsons-symmetric",
                          "--cells_per_mfp", "10", "--num_angles", "4", "--steps", "20"], capture_output=True, text=True,
                         cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr          # the reference crashes here (defects D1-D5)
    assert "alphas =" in out.stdout


def test_cli_extras_subcritical_variant_and_output(tmp_path):
    """--nu_sigmaf / --dt / --output: the sub-critical K-P slab of the report's Table 1
    (nu*Sigma_f = 0.3, k = 0.4243 on a fine mesh) through the CLI, against the oracle on the same mesh."""
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem
    out_file = str(tmp_path / "flux.npz")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--problem", "kornreich-parsons-symmetric",
                          "--cells_per_mfp", "50", "--num_angles", "16", "--nu_sigmaf", "0.3", "--output", out_file],
                         capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    p = kp_problem(9, 50, 16, nusf_fuel=0.3)
    ro = orc.power_iteration(p["mu"], p["weight"], p["hx"], p["tables"], p["phi0"])
    got = np.load(out_file)
    assert abs(float(got["k"]) - ro["k_new"]) < 1e-9 and abs(float(got["k"]) - 0.4243178) < 1e-3     # report: fine mesh
    assert _rel(got["phi"], ro["phi_new"]) < 1e-10 and got["midpoints"].shape == (450,)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--problem", "kornreich-parsons-symmetric",
                          "--cells_per_mfp", "10", "--num_angles", "4", "--steps", "20", "--dt", "0.05", "--nu_sigmaf", "0.3",
                          "--output", out_file], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr
    got = np.load(out_file)
    assert float(got["dt"]) == 0.05 and got["phi_steps"].shape == (90, 1, 20) and np.all(np.real(got["alphas"]) < 0)


def test_sweep1D_honours_explicit_sigT_in_both_directions():
    """Repair R5: the reference ignores sigT for mu < 0 (group.py:56); the oracle and the device do not."""
    from oracle import ae_oracle as orc
    s, g = _kp(10, 4)
    rs = np.random.RandomState(1)
    src = rs.rand(s.num_zones, 1)
    sigT = g.total_xs + 10.0 * g.inv_speed
    for n in range(4):
        psi = g.sweep1D(g.angle[n], src, sigT)
        assert _rel(psi, orc.sweep(s.mu[n], s.hx, src, sigT)) < 1e-13


def test_one_step_flux_matches_oracle_step():
    """TimeUnit.calculate_one_step_flux with an explicit per-angle source (time_unit.py:29-42)."""
    from alpha_eigen_modules.time_unit import TimeUnit
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem, kp_psi0
    s, g = _kp(10, 4)
    tu = TimeUnit(slab=s, group=g, num_steps=1, dt=0.1)
    p = kp_problem(9, 10, 4)
    psi0 = kp_psi0(p)
    source = psi0 * (p["inv_speed"][:, None, :] / 0.1)                  # S + psi_prev*inv_speed/dt with S = 0
    outer, inner = tu.calculate_one_step_flux(source)
    res = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)), psi0, 0.1, 1)
    assert (outer, inner) == (int(res["outer"][0]), int(res["inner"][0]))
    assert _rel(g.phi.new, res["phi_last"]) < 1e-10
    psi = np.stack([a.psi_midpoint for a in g.angle], axis=1)
    assert _rel(psi, res["psi_last"]) < 1e-10


def test_time_dependent_source_iteration_method():
    """TimeUnit.time_dependent_source_iteration(sigT, source): inner scatter iteration for an explicit Q."""
    from alpha_eigen_modules.time_unit import TimeUnit
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem
    s, g = _kp(10, 4)
    tu = TimeUnit(slab=s, group=g, num_steps=1, dt=0.1)
    p = kp_problem(9, 10, 4)
    rs = np.random.RandomState(3)
    Q = rs.rand(p["Z"], 4, 1) + 0.1
    iters = tu.time_dependent_source_iteration(None, Q)
    # oracle: the same loop written out with single sweeps (time_unit.py:47-58)
    sigT = p["total_xs"] + 1 / 0.1 * p["inv_speed"]
    phi_old = np.zeros((p["Z"], 1))
    for it in range(1, 100):
        phi = np.zeros((p["Z"], 1))
        for n in range(4):
            phi[:, 0] += orc.sweep(p["mu"][n], p["hx"], Q[:, n] + phi_old * p["scatter_xs"] * 0.5, sigT) * p["weight"][n]
        change = np.linalg.norm(phi - phi_old) / np.linalg.norm(phi)
        phi_old = phi
        if change < 1e-10:
            break
    assert iters == it
    assert _rel(g.phi.new, phi_old) < 1e-10
    psi_last = np.stack([a.psi_midpoint[:, 0] for a in g.angle], axis=1)
    assert psi_last.shape == (p["Z"], 4) and np.all(np.isfinite(psi_last))


def test_multigroup_classes_against_oracle():
    """MultiGroup + TableMaterial on the scaled five-region slab: k and a short time march."""
    from alpha_eigen_b200.group import MultiGroup
    from alpha_eigen_b200.slab import ScaledHeterogeneousSlab
    from alpha_eigen_b200.synthetic import synthetic_materials, synthetic_problem
    from alpha_eigen_b200.time_unit import TimeUnit
    from oracle import ae_oracle as orc
    Z, A, G = 700, 8, 12
    fuel, mod, absb = synthetic_materials(G)
    np.random.seed(2021)
    s = ScaledHeterogeneousSlab(Z, A, mod, fuel, absb)
    assert s.num_groups == G
    mg = MultiGroup(s)
    mg.create_zones()
    p = synthetic_problem(Z, A, G)
    np.testing.assert_array_equal(mg._mat_index == 0, p["mat"] == 0)    # same region layout as the array builder
    tab = orc.Tables(mg._mat_index, *[np.array([getattr(m, n) for m in mg._materials]) for n in
                                      ("total_xs", "scatter_xs", "chi", "nu_sigmaf", "inv_speed")])
    phi0 = mg.phi.old.copy()
    k = mg.perform_power_iteration(quiet=True)
    ro = orc.power_iteration(s.mu, s.weight, s.hx, tab, phi0)
    assert ro["status"] == 0 and abs(k.new - ro["k_new"]) < 1e-9
    assert _rel(mg.phi.new, ro["phi_new"]) < 1e-10
    steps = 3
    psi0 = np.repeat(0.5 * mg.phi.new[:, None, :], A, axis=1)
    tu = TimeUnit(slab=s, group=mg, num_steps=steps, dt=0.1)
    mg.source = np.zeros((Z, G))
    tu.calculate_time_dependent_flux()
    res = orc.time_march(s.mu, s.weight, s.hx, tab, np.zeros((Z, G)), psi0, 0.1, steps)
    assert list(tu.outer_iterations) == list(res["outer"]) and list(tu.inner_iterations) == list(res["inner"])
    assert _rel(tu.psi, res["psi_snap"]) < 1e-10
    assert _rel(tu.phi, res["phi_snap"]) < 1e-10


def test_pack_unpack_roundtrip_and_padding():
    from alpha_eigen_b200.device import SlabDevice
    Z, A, G = 1000, 6, 3
    mu, w = np.polynomial.legendre.leggauss(A)
    one = np.ones((1, G))
    dev = SlabDevice(Z, 1.0, mu, w, np.zeros(Z, dtype=np.int64), one, np.zeros((1, G, G)), one, one, one)
    rs = np.random.RandomState(0)
    a = rs.rand(Z, G)
    t = dev.cells_to_device(a)
    np.testing.assert_array_equal(dev.cells_to_host(t), a)
    raw = t.cpu().numpy().reshape(G, dev.Zp)
    assert dev.Zp == 1024 and np.count_nonzero(raw) == Z * G            # padding cells are zero
    psi = rs.rand(Z, A, G)
    tp = dev.psi_to_device(psi)
    np.testing.assert_array_equal(dev.psi_to_host(tp), psi[:, dev.angles, :])


def test_flux_update_entry_point_matches_numpy():
    """ae_sweep_set + ae_flux_update as bench.py chains them: phi = sum of partials, change, next source."""
    from oracle import ae_oracle as orc
    from tests.test_gpu_parity import _device, _mg_problem
    Z, A, G = 900, 16, 12
    p = _mg_problem(Z, A, G, seed=5)
    t = p["tables"]
    dev = _device(p)
    q = p["rs"].rand(Z, G)
    base = p["rs"].rand(Z, G)
    phi_prev = p["rs"].rand(Z, G)
    dev.base.copy_(dev.cells_to_device(base))
    dev.phi.copy_(dev.cells_to_device(phi_prev))
    dq = dev.cells_to_device(q)
    import ctypes as C
    from alpha_eigen_b200 import _lib
    _lib.check(dev.lib.ae_sweep_set(C.byref(dev.slab), dq.data_ptr(), None, None, dev.partial.data_ptr(),
                                    dev.sync.data_ptr(), dev.stream))
    dev.flux_update(dev.partial, dev.P, dev.q1, tol=1e-30, iteration=7)
    phi_o = orc.sweep_set(p["mu"], p["weight"], p["hx"], q, t.total[t.mat])
    assert _rel(dev.cells_to_host(dev.phi), phi_o) < 1e-13
    q_next = base + 0.5 * np.einsum("zp,zpg->zg", phi_o, t.scat[t.mat])
    assert _rel(dev.cells_to_host(dev.q1), q_next) < 1e-13
    ctrl = dev.read_ctrl()
    change = np.linalg.norm(phi_o - phi_prev) / np.linalg.norm(phi_o)
    assert abs(ctrl.change - change) < 1e-12 * change and ctrl.inner_iters == 7 and ctrl.inner_done == 0


@pytest.mark.parametrize("mode", ["batched", "streams"])
def test_ensemble_members_match_standalone_oracle(mode):
    """BASELINE configs[4] in miniature: every member of a concurrently solved ensemble equals its own
    stand-alone oracle solve (iteration counts, psi) and carries its own alpha spectrum.  batched = all members
    in one cooperative kernel (ae_ensemble_time_march), streams = one fused kernel per member."""
    from alpha_eigen_b200.ensemble import Ensemble, synthetic_ensemble
    from oracle import ae_oracle as orc
    E, Z, A, G, steps = 5, 300, 4, 3, 12
    members = synthetic_ensemble(E, Z, A, G)
    res = Ensemble(members, dt=0.1, num_steps=steps, concurrent=3, mode=mode).run(keep_snapshots=True)
    dom = []
    for p, r in zip(members, res):
        tab = orc.Tables(p["mat"], p["total"], p["scat"], p["chi"], p["nusf"], p["invspeed"])
        psi0 = np.repeat(0.5 * p["phi0"][:, None, :], A, axis=1)
        ro = orc.time_march(p["mu"], p["weight"], p["hx"], tab, np.zeros((Z, G)), psi0, 0.1, steps, want_phi_snap=False)
        assert r["status"] == 0 and ro["status"] == 0
        assert list(r["outer"]) == list(ro["outer"]) and list(r["inner"]) == list(ro["inner"])
        assert _rel(r["psi_last"], ro["psi_last"][:, r["angles"], :]) < 1e-10
        ref = orc.dmd_alpha(ro["psi_snap"].reshape(-1, steps), steps, 0.1)
        ref = ref[np.lexsort((ref.imag, -ref.real))]
        assert r["rank"] == len(ref)
        assert abs(r["alphas"][0] - ref[0]) < 1e-6 * abs(ref[0])
        dom.append(r["alphas"][0].real)
    assert len(set(np.round(dom, 8))) == E          # parameter variation shows up in the spectrum


def test_batched_ensemble_queue_and_failed_member():
    """More members than teams fit on the chip (the device-side queue hands them out), members of a shape with two
    angle blocks per direction, and one member that cannot converge: it must be reported alone, the others must
    equal the one-kernel-per-member path bit for bit (same arithmetic, same order)."""
    import torch
    from alpha_eigen_b200 import _lib
    from alpha_eigen_b200.device import SlabDevice
    from alpha_eigen_b200.ensemble import march_batched, synthetic_ensemble
    E, Z, A, G, steps = 70, 520, 20, 4, 4          # S20: 10 angles per direction = 2 blocks of 8 -> 16 CTAs per member
    members = synthetic_ensemble(E, Z, A, G)
    bad = 37
    members[bad]["nusf"] = members[bad]["nusf"] * 400.0        # wildly supercritical: the fission loop diverges
    devs, psis, snaps, phis = [], [], [], []
    for p in members:
        dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                         p["invspeed"], dt=0.1)
        devs.append(dev)
        psis.append(dev.psi_to_device(np.repeat(0.5 * p["phi0"][:, None, :], A, axis=1)))
        snaps.append(torch.zeros(steps * dev.G * dev.A * dev.Zp, dtype=torch.float64, device=dev.dev))
        phis.append(torch.zeros(steps * dev.G * dev.Zp, dtype=torch.float64, device=dev.dev))
    status, outer, inner, teams = march_batched(devs, psis, snaps, phis, steps)
    assert 1 <= teams < E, teams
    assert status[bad] == _lib.AE_NOT_CONVERGED and np.count_nonzero(status) == 1
    for j in (0, 1, bad - 1, bad + 1, E - 1):
        p = members[j]
        solo = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                          p["invspeed"], dt=0.1)
        psi = solo.psi_to_device(np.repeat(0.5 * p["phi0"][:, None, :], A, axis=1))
        snap = torch.zeros_like(snaps[j])
        phi_snap = torch.zeros_like(phis[j])
        st, o, i = solo.time_march(psi, steps, 1e-10, snap, phi_snap)
        assert st == 0 and list(o) == list(outer[j]) and list(i) == list(inner[j])
        assert bool((snap == snaps[j]).all()) and bool((phi_snap == phis[j]).all()) and bool((psi == psis[j]).all())
        assert bool((solo.phi == devs[j].phi).all())


def test_ensemble_rejects_mixed_shapes():
    import torch
    from alpha_eigen_b200.device import SlabDevice
    from alpha_eigen_b200.ensemble import march_batched, synthetic_ensemble
    a = synthetic_ensemble(1, 300, 4, 3)[0]
    b = synthetic_ensemble(1, 600, 4, 3)[0]
    devs, psis = [], []
    for p in (a, b):
        dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                         p["invspeed"], dt=0.1)
        devs.append(dev)
        psis.append(dev.new_psi())
    with pytest.raises(ValueError, match="different shape"):
        march_batched(devs, psis, None, None, 2)


def test_multi_region_slab_generic_zone_path():
    """MultiRegionSlab (slab.py:59-68, repaired) + Zone.assign_material (zone.py:16-36): the per-zone Python
    path of create_zones (no vectorised map) feeding the device, against the oracle."""
    from alpha_eigen_b200.group import MultiGroup
    from alpha_eigen_b200.slab import MultiRegionSlab
    from alpha_eigen_b200.synthetic import synthetic_materials
    from oracle import ae_oracle as orc
    G = 4
    metal, poly, _ = synthetic_materials(G, seed=5)
    np.random.seed(2021)
    s = MultiRegionSlab(length=10, cells_per_mfp=1, num_angles=6, reflector_thick=2, inner_thick=2, num_zones=333,
                        num_groups=G, material1=metal, material2=poly)
    mg = MultiGroup(s)
    mg.create_zones()
    names = [z.region_name for z in mg.zone]
    assert names[0] == "reflector" and names[166] == "moderator" and "fuel" in names and names[-1] == "reflector"
    assert mg.zone[0].out_group_scatter_xs.shape == (G, G) and np.all(np.diag(mg.zone[0].out_group_scatter_xs) == 0)
    tab = orc.Tables(mg._mat_index, *[np.array([getattr(m, n) for m in mg._materials]) for n in
                                      ("total_xs", "scatter_xs", "chi", "nu_sigmaf", "inv_speed")])
    phi0 = mg.phi.old.copy()
    k = mg.perform_power_iteration(quiet=True)
    ro = orc.power_iteration(s.mu, s.weight, s.hx, tab, phi0)
    assert ro["status"] == 0 and abs(k.new - ro["k_new"]) < 1e-9
    assert _rel(mg.phi.new, ro["phi_new"]) < 1e-10


def test_host_pipe_overlapped_io_roundtrip():
    """HostPipe: double-buffered upload / pack and unpack / download on side streams; five steps with distinct
    sources, each result must be the source of its own step (slot reuse is ordered by events)."""
    import torch
    from alpha_eigen_b200.device import HostPipe, SlabDevice
    Z, A, G = 3000, 4, 6
    mu, w = np.polynomial.legendre.leggauss(A)
    one = np.ones((1, G))
    dev = SlabDevice(Z, 1.0, mu, w, np.zeros(Z, dtype=np.int64), one, np.zeros((1, G, G)), one, one, one)
    pipe = HostPipe(dev)
    n = 5
    src = [torch.rand(Z, G, dtype=torch.float64).pin_memory() for _ in range(n)]
    out = [torch.zeros(Z, G, dtype=torch.float64).pin_memory() for _ in range(n)]
    rows = [torch.zeros(G * dev.Zp, dtype=torch.float64, device=dev.dev) for _ in range(2)]
    pipe.upload(0, src[0])
    for i in range(n):
        pipe.feed(i, rows[i & 1])
        if i + 1 < n:
            pipe.upload(i + 1, src[i + 1])
        rows[i & 1].mul_(2.0)                                       # the "step"
        pipe.drain(i, rows[i & 1], out[i])
    pipe.join()
    torch.cuda.current_stream().synchronize()
    pipe.wait()
    for i in range(n):
        np.testing.assert_array_equal(out[i].numpy(), 2.0 * src[i].numpy())


def test_real_70_group_data_through_the_public_classes():
    """the reference's own 70-group metal / polyethylene cross-sections
    (material.py:26-46; packaged copy, read through Material) on a reflected metal slab: k power iteration and
    a short time march through MultiGroup / TimeUnit against the oracle on the same tables."""
    from alpha_eigen_b200.group import MultiGroup
    from alpha_eigen_b200.material import Material
    from alpha_eigen_b200.slab import MultiRegionSlab
    from alpha_eigen_b200.time_unit import TimeUnit
    from oracle import ae_oracle as orc
    G = 70
    metal = Material("metal", G, use_file_chi=True)         # material.py:43 zeroes chi: no fission source at all
    poly = Material("polyethylene", G, use_file_chi=True)
    np.random.seed(2021)
    s = MultiRegionSlab(length=10, cells_per_mfp=1, num_angles=8, reflector_thick=2, inner_thick=1, num_zones=200,
                        num_groups=G, material1=metal, material2=poly)
    mg = MultiGroup(s)
    mg.create_zones()
    tab = orc.Tables(mg._mat_index, *[np.array([getattr(m, n) for m in mg._materials]) for n in
                                      ("total_xs", "scatter_xs", "chi", "nu_sigmaf", "inv_speed")])
    phi0 = mg.phi.old.copy()
    k = mg.perform_power_iteration(quiet=True)
    ro = orc.power_iteration(s.mu, s.weight, s.hx, tab, phi0)
    assert ro["status"] == 0 and k.converged
    assert abs(k.new - ro["k_new"]) < 1e-9 * ro["k_new"], (k.new, ro["k_new"])
    assert _rel(mg.phi.new, ro["phi_new"]) < 1e-10
    # two backward-Euler steps from the converged flux shape (metal inv_speed spans 1.8e-2 .. 4.3e3); the fixed
    # source of the march is group.source (time_unit.py:19), which the power iteration left as the fission source
    steps = 2
    tu = TimeUnit(s, mg, steps, 0.1)
    psi0 = np.repeat(0.5 * mg.phi.new[:, None, :], s.num_angles, axis=1)
    tu.psi0 = psi0
    s_ext = np.asarray(mg.source, dtype=np.float64).reshape(s.num_zones, G).copy()
    tu.calculate_time_dependent_flux()
    rt = orc.time_march(s.mu, s.weight, s.hx, tab, s_ext, psi0, 0.1, steps)
    assert rt["status"] == 0
    assert list(tu.outer_iterations) == list(rt["outer"]) and list(tu.inner_iterations) == list(rt["inner"])
    assert _rel(tu.psi[..., steps - 1], rt["psi_last"]) < 1e-10


def test_cell_range_host_io_roundtrip():
    """ae_pack_cells_range / ae_unpack_cells_range (single rank: the range is the whole row)."""
    import torch
    from alpha_eigen_b200.device import SlabDevice
    Z, A, G = 700, 4, 5
    mu, w = np.polynomial.legendre.leggauss(A)
    one = np.ones((1, G))
    dev = SlabDevice(Z, 1.0, mu, w, np.zeros(Z, dtype=np.int64), one, np.zeros((1, G, G)), one, one, one)
    z0, count, s0, cells = dev.cell_range()
    assert (z0, count, s0, cells) == (0, Z, 0, dev.Zp)
    src = torch.rand(Z, G, dtype=torch.float64).pin_memory()
    rows = torch.zeros(G * dev.Zp, dtype=torch.float64, device=dev.dev)
    dev.shard_to_device(src, rows)
    np.testing.assert_array_equal(dev.cells_to_host(rows), src.numpy())
    back = torch.empty(Z, G, dtype=torch.float64).pin_memory()
    dev.shard_to_host(rows, back)
    np.testing.assert_array_equal(back.numpy(), src.numpy())


def test_device_problem_is_rebuilt_when_the_source_moves():
    """OneGroup keeps the problem resident between calls; the cache key is a digest of every input byte, so a point
    source moved to another cell (same sum, same sum of squares) must give the other answer, not the cached one."""
    from alpha_eigen_modules.time_unit import TimeUnit
    out = []
    for cell in (30, 61):
        s, g = _kp(10, 4)
        tu = TimeUnit(slab=s, group=g, num_steps=2, dt=0.1)
        g.source = np.zeros((s.num_zones, 1))
        g.source[30, 0] = 1.0
        tu.calculate_time_dependent_flux()                  # builds and caches the device problem
        g.source = np.zeros((s.num_zones, 1))
        g.source[cell, 0] = 1.0                             # same shape, sum and sum of squares
        tu2 = TimeUnit(slab=s, group=g, num_steps=2, dt=0.1)
        tu2.calculate_time_dependent_flux()
        out.append(tu2.phi[:, 0, -1].copy())
    assert not np.allclose(out[0], out[1])
    assert int(np.argmax(out[1] - out[0])) in range(55, 68)     # the extra flux sits where the source went


def test_explicit_sigT_must_match_the_device_problem():
    """calculate_one_step_flux / time_dependent_source_iteration take sigT like the reference (time_unit.py:29,44); the
    device problem already carries total_xs + inv_speed/dt, so any other array is refused instead of being ignored."""
    from alpha_eigen_modules.time_unit import TimeUnit
    s, g = _kp(10, 4)
    tu = TimeUnit(slab=s, group=g, num_steps=1, dt=0.1)
    src = np.ones((s.num_zones, s.num_angles, 1))
    good = g.total_xs + g.inv_speed / 0.1
    tu.calculate_one_step_flux(src, sigT=good)
    with pytest.raises(ValueError):
        tu.calculate_one_step_flux(src, sigT=good * 1.01)
    g.inv_speed[5, 0] = 0.0
    with pytest.raises(ValueError):
        tu.time_dependent_source_iteration(sigT=None, source=src)
