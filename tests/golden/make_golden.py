"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) in this container
(the `march` mode drives the reference's own sweep and inner iteration under the documented minimal repairs of
its broken time loop, applied from outside -- see the comment there).

The reference cannot travel to the GPU box, so its outputs are committed as small .npz fixtures
next to this script.  Run from the repo root:

    python tests/golden/make_golden.py

Every configuration runs in a FRESH interpreter because the reference seeds the global NumPy RNG
at import (phi.py:2) and consumes it in OneGroup.__init__ / Zone.__init__ (group.py:18, zone.py:11);
the reference's own goldens were produced one configuration per process.
Nothing is written under /root/reference (PYTHONDONTWRITEBYTECODE=1, h5py stub in /tmp).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = "/tmp/ae_golden_stubs"

WORKER = r'''
import sys, warnings, json
warnings.filterwarnings("ignore")
import numpy as np
mode, out = sys.argv[1], sys.argv[2]
from alpha_eigen_modules.slab import SymmetricHeterogeneousSlab
from alpha_eigen_modules.group import OneGroup
from alpha_eigen_modules.material import SimpleMaterial
from alpha_eigen_modules.time_unit import TimeUnit

def build(length, cpm, A):
    fuel = SimpleMaterial('fuel', num_groups=1)
    moderator = SimpleMaterial('moderator', num_groups=1)
    absorber = SimpleMaterial('absorber', num_groups=1)
    slab = SymmetricHeterogeneousSlab(length, cpm, A, moderator, fuel_mat=fuel, abs_mat=absorber)
    g = OneGroup(slab=slab)
    g.create_zones()
    return slab, g

if mode == "k":
    length, cpm, A = map(int, sys.argv[3:6])
    slab, g = build(length, cpm, A)
    phi0_new, phi0_old = g.phi.new.copy(), g.phi.old.copy()
    counts = {"sweeps": 0, "si": []}
    orig_si = OneGroup.source_iteration_for_scalar_flux
    orig_sw = OneGroup.sweep1D
    def sw(self, angle, source, sigT=None):
        counts["sweeps"] += 1
        return orig_sw(self, angle, source, sigT)
    def si(self, *a, **k):
        before = counts["sweeps"]
        r = orig_si(self, *a, **k)
        counts["si"].append((counts["sweeps"] - before) // self.slab.num_angles)
        return r
    OneGroup.sweep1D = sw
    OneGroup.source_iteration_for_scalar_flux = si
    k = g.perform_power_iteration(quiet=True)
    psi = np.stack([a.psi_midpoint[:, 0] for a in g.angle], axis=1)   # (Z, A)
    np.savez_compressed(out, length=length, cells_per_mfp=cpm, num_angles=A,
             k_new=k.new, k_old=k.old, phi0=phi0_old, phi_new=g.phi.new, phi_old=g.phi.old,
             psi_midpoint=psi, si_hist=np.array(counts["si"], dtype=np.int32),
             total_xs=g.total_xs, scatter_xs=g.scatter_xs, chi=g.chi, nu_sigmaf=g.nu_sigmaf,
             mu=slab.mu, weight=slab.weight, hx=slab.hx)
elif mode == "sweep":
    # single sweeps + one source iteration on the K-P slab with a seeded arbitrary source
    length, cpm, A = map(int, sys.argv[3:6])
    slab, g = build(length, cpm, A)
    rs = np.random.RandomState(7)
    src = rs.rand(slab.num_zones, 1) + 0.1
    psis = np.stack([g.sweep1D(g.angle[n], src) for n in range(A)], axis=1)       # (Z, A)
    g.source = src.copy()
    g.source_iteration_for_scalar_flux()
    psi_after = np.stack([a.psi_midpoint[:, 0] for a in g.angle], axis=1)
    np.savez_compressed(out, length=length, cells_per_mfp=cpm, num_angles=A, source=src, psi=psis,
             phi_si=g.phi.new, psi_si=psi_after, total_xs=g.total_xs, scatter_xs=g.scatter_xs,
             mu=slab.mu, weight=slab.weight, hx=slab.hx)
elif mode == "zone":
    # Multigroup source formulas of the reference, the only executable multigroup code it has: Zone.compute_fission_source
    # and Zone.compute_scattering_source (zone.py:38-51) on seeded fluxes and cross-sections.
    from alpha_eigen_modules.zone import Zone
    G, Z = int(sys.argv[3]), int(sys.argv[4])
    rs = np.random.RandomState(5)
    class S: pass
    slab = S(); slab.num_groups = G
    class M: pass
    mats = []
    for _ in range(3):
        m = M()
        m.scatter_xs = rs.rand(G, G)            # [g_from, g_to] (material.py:38)
        m.nu_sigmaf = rs.rand(G)
        m.chi = rs.rand(G); m.chi /= m.chi.sum()
        mats.append(m)
    mat_idx = rs.randint(0, 3, Z)
    phi = rs.rand(Z, G) + 0.1
    fission = np.zeros((Z, G)); source = np.zeros((Z, G))
    for i in range(Z):
        z = Zone(slab, float(i))
        z.material = mats[mat_idx[i]]
        z.out_group_scatter_xs = z.material.scatter_xs.copy()          # zone.py:33-36
        for g in range(G):
            z.out_group_scatter_xs[g, g] = 0
        z.phi.old = phi[i].copy(); z.phi.new = phi[i].copy()
        z.compute_fission_source()                                      # zone.py:38-42
        fission[i] = z.fission_source
        z.source = np.zeros(G)
        z.compute_scattering_source()                                   # zone.py:48-51 (out-group scatter + fission)
        source[i] = z.source
    np.savez_compressed(out, G=G, Z=Z, mat=mat_idx, phi=phi, scat=np.array([m.scatter_xs for m in mats]),
             nusf=np.array([m.nu_sigmaf for m in mats]), chi=np.array([m.chi for m in mats]),
             fission_source=fission, scatter_plus_fission_source=source)
elif mode == "march":
    # Time march with the reference's OWN sweep1D (group.py:41-60) and time_dependent_source_iteration
    # (time_unit.py:44-59), both unmodified, under the minimal repairs of SURVEY section 3.3 applied from outside:
    #   R3  tu.phi becomes an ndarray SUBCLASS, so `self.phi.new = phi_old` (time_unit.py:59) sticks instead of raising;
    #   R5  every zone gets a material shim whose total_xs is the time-augmented sigT of that zone, which is what the
    #       negative-mu branch reads (group.py:56);
    #   R2 / R4  the two outer loops that cannot work as written (time_unit.py:17-27 reads the psi slot it is about to
    #       fill, :30-42 never updates phi) are driven here, line by line as the reference spells them, with the
    #       previous step's psi and the inner solve's flux.
    # psi0 = phi0 / 2 for every angle (DESIGN.md section 2), phi0 the seeded symmetric guess of phi.py:10-11.
    length, cpm, A, steps = map(int, sys.argv[3:7])
    dt = 0.1                                                            # inputs.py:39
    slab, g = build(length, cpm, A)
    Z = slab.num_zones
    tu = TimeUnit(slab, g, steps, dt)
    class PhiArray(np.ndarray): pass
    tu.phi = tu.phi.view(PhiArray)                                      # R3
    sigT = g.total_xs + 1 / dt * g.inv_speed                            # time_unit.py:24
    class Shim: pass
    for i in range(Z):                                                  # R5
        m = Shim(); m.__dict__.update(g.zone[i].material.__dict__); m.total_xs = float(sigT[i, 0])
        g.zone[i].material = m
    phi0 = g.phi.new.copy()
    psi_prev = np.repeat(0.5 * phi0[:, None, :], A, axis=1)             # (Z, A, 1)
    for n in range(A):
        tu.source[:, n] = g.source[:]                                   # time_unit.py:18-19 (zero: no power iteration ran)
    outer_hist, inner_hist = [], []
    calls = {"n": 0}
    orig = OneGroup.sweep1D
    def counted(self, angle, source, sigT=None):
        calls["n"] += 1
        return orig(self, angle, source, sigT)
    OneGroup.sweep1D = counted
    for step in range(steps):
        source = tu.source.copy()
        for n in range(A):
            source[:, n, 0] = tu.source[:, n, 0] + psi_prev[:, n, 0] * g.inv_speed[:, 0] / dt     # :23 with R2
        phi = np.zeros((Z, 1))                                          # :30
        iteration, before = 0, calls["n"]
        while True:
            iteration += 1
            assert iteration < 100
            phi_old = phi.copy()                                        # :36
            Q = source.copy()
            for n in range(A):
                Q[:, n, 0] += 0.5 * (g.chi[:, 0] * phi[:, 0] * g.nu_sigmaf[:, 0])                 # :38-39
            tu.time_dependent_source_iteration(sigT, Q)                 # :40, unmodified reference code
            phi = np.asarray(tu.phi.new).copy()                         # R4
            change = np.linalg.norm(np.reshape(phi - phi_old, (Z, 1))) / np.linalg.norm(np.reshape(phi, (Z, 1)))   # :41
            if change < 1e-10:
                break
        outer_hist.append(iteration)
        inner_hist.append((calls["n"] - before) // A)
        tu.phi[:, 0, step] = np.reshape(phi, Z)                         # :25
        for n in range(A):
            tu.psi[:, n, :, step] = g.angle[n].psi_midpoint             # :26-27
        psi_prev = tu.psi[:, :, :, step].copy()
    np.savez_compressed(out, length=length, cells_per_mfp=cpm, num_angles=A, steps=steps, dt=dt, phi0=phi0,
             psi=np.asarray(tu.psi), phi=np.asarray(tu.phi), outer=np.array(outer_hist, dtype=np.int32),
             inner=np.array(inner_hist, dtype=np.int32))
elif mode == "dmd":
    # unmodified TimeUnit.compute_alpha_eigenvalues on seeded snapshot arrays (SURVEY appendix A)
    kind, Z, A, steps = sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    dt = 0.1
    class S: pass
    s = S(); s.num_zones, s.num_angles, s.num_groups = Z, A, 1
    tu = TimeUnit(s, None, steps, dt)
    rs = np.random.RandomState(11)
    if kind == "modes":
        # psi(t) = sum_j c_j v_j / (1 - alpha_j dt)^step : exact backward-Euler modes + decaying noise floor
        alphas = np.array([-0.05, -0.3, -1.0, -2.5, 0.2])
        V = rs.rand(Z * A, len(alphas)) + 0.1
        lam = 1.0 / (1.0 - alphas * dt)
        for st in range(steps):
            col = V @ (np.array([1.0, 0.7, 0.5, 0.3, 1e-3]) * lam ** (st + 1))
            tu.psi[:, :, 0, st] = col.reshape(Z, A)
    else:
        snap = np.load(sys.argv[7])["psi_snap"]           # (Z, A, 1, steps) from the repaired oracle march (pinned bit
                                                          # for bit to the reference's own sweeps by kp_march_repaired_*)
        tu.psi[...] = snap
    al = tu.compute_alpha_eigenvalues()
    if kind == "march_small":
        # large march: keep the answers, not the 11 MB of snapshots (the march itself is pinned elsewhere)
        first = 2 * int(steps / 5)
        K = int(steps - int(steps / 5) - steps / 50) - 1 - int(steps / 5)
        sv = np.linalg.svd(np.asarray(tu.psi).reshape(Z * A, steps)[:, first:first + K], compute_uv=False)
        np.savez_compressed(out, Z=Z, A=A, steps=steps, dt=dt, alphas=np.asarray(al, dtype=complex), sigma=sv,
                            psi_last=np.asarray(tu.psi)[..., -1], phi_last=np.load(sys.argv[7])["phi_last"],
                            outer=np.load(sys.argv[7])["outer"], inner=np.load(sys.argv[7])["inner"])
    else:
        np.savez_compressed(out, Z=Z, A=A, steps=steps, dt=dt, psi=tu.psi, alphas=np.asarray(al, dtype=complex))
'''


def run(args):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=f"{STUBS}:/root/reference")
    subprocess.check_call([sys.executable, "-W", "ignore", "-c", WORKER] + [str(a) for a in args],
                          env=env, cwd="/tmp")


def main():
    os.makedirs(STUBS, exist_ok=True)
    with open(os.path.join(STUBS, "h5py.py"), "w") as f:        # material.py:2 imports h5py unconditionally
        f.write('class File:\n    def __init__(self,*a,**k): raise RuntimeError("h5py stub")\n')
    for length, cpm, A in [(9, 5, 2), (9, 10, 4), (9, 20, 8), (9, 50, 4), (9, 100, 16)]:
        out = os.path.join(HERE, f"kp_k_{length}_{cpm}_{A}.npz")
        print("k golden", length, cpm, A, flush=True)
        run(["k", out, length, cpm, A])
    for length, cpm, A in [(9, 10, 4), (9, 37, 6)]:
        out = os.path.join(HERE, f"kp_sweep_{length}_{cpm}_{A}.npz")
        print("sweep golden", length, cpm, A, flush=True)
        run(["sweep", out, length, cpm, A])
    for length, cpm, A, steps in [(9, 5, 2, 6), (9, 10, 4, 4)]:
        out = os.path.join(HERE, f"kp_march_repaired_{length}_{cpm}_{A}_{steps}.npz")
        print("march golden (reference sweep + inner iteration, repaired outer loops)", length, cpm, A, steps, flush=True)
        run(["march", out, length, cpm, A, steps])
    print("zone source formulas golden", flush=True)
    run(["zone", os.path.join(HERE, "mg_zone_sources_7_40.npz"), 7, 40])
    print("dmd golden (modes)", flush=True)
    run(["dmd", os.path.join(HERE, "dmd_modes_64_4_50.npz"), "modes", 64, 4, 50])
    run(["dmd", os.path.join(HERE, "dmd_modes_40_2_200.npz"), "modes", 40, 2, 200])
    # DMD of a real (repaired) time march: snapshots come from the oracle, alphas from the reference routine
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests.kp import kp_problem, kp_psi0
    from oracle import ae_oracle as orc
    import numpy as np
    for cpm, A, steps in [(10, 4, 50)]:
        p = kp_problem(9, cpm, A)
        res = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)),
                             kp_psi0(p), 0.1, steps)
        assert res["status"] == 0
        tmp = f"/tmp/ae_march_{cpm}_{A}_{steps}.npz"
        np.savez(tmp, psi_snap=res["psi_snap"])
        print("dmd golden (march)", cpm, A, steps, flush=True)
        run(["dmd", os.path.join(HERE, f"dmd_march_{cpm}_{A}_{steps}.npz"), "march", p["Z"], A, steps, tmp])
    # the well-conditioned real march (the reference routine is reproducible to ~1e-9 on it): answers only
    for cpm, A, steps in [(50, 16, 200)]:
        p = kp_problem(9, cpm, A)
        res = orc.time_march(p["mu"], p["weight"], p["hx"], p["tables"], np.zeros((p["Z"], 1)),
                             kp_psi0(p), 0.1, steps)
        assert res["status"] == 0
        tmp = f"/tmp/ae_march_{cpm}_{A}_{steps}.npz"
        np.savez(tmp, psi_snap=res["psi_snap"], phi_last=res["phi_last"], outer=res["outer"], inner=res["inner"])
        print("dmd golden (march, answers only)", cpm, A, steps, flush=True)
        run(["dmd", os.path.join(HERE, f"dmd_march_answers_{cpm}_{A}_{steps}.npz"), "march_small", p["Z"], A, steps, tmp])


if __name__ == "__main__":
    main()
