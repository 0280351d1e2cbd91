"""ctypes front end of oracle/ae_oracle.c plus the NumPy restatement of the DMD step.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by alpha_eigen_b200/.

Parity status (details in ae_oracle.c's header and DESIGN.md):
  * one-group sweep / source iteration / power iteration: pinned by the reference's three
    golden k values and by phi/psi vectors produced by the unmodified reference
    (tests/golden/make_golden.py).
  * compute_alpha_eigenvalues (time_unit.py:61-92): pinned by outputs of the unmodified
    reference routine on seeded snapshot arrays (tests/golden/dmd_*.npz).
  * time march (one group): the reference's loop crashes as shipped; pinned against snapshots produced by the
    reference's unmodified sweep1D and time_dependent_source_iteration under the minimal repairs R2-R5 applied from
    outside (tests/golden/kp_march_repaired_*.npz, make_golden.py mode `march`): bit-exact.
  * G > 1: the fission and scatter source terms are pinned against the reference's Zone.compute_fission_source /
    compute_scattering_source (zone.py:38-51; tests/golden/mg_zone_sources_7_40.npz); the iteration around them is
    PARITY UNPINNED (the reference has no runnable multigroup driver).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libae_oracle.so")
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile ae_oracle.c with gcc (-ffp-contract=off keeps NumPy's one-rounding-per-op)."""
    src = os.path.join(_HERE, "ae_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libae_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.aeo_sweep.restype = None
        _lib.aeo_sweep_set.restype = None
        _lib.aeo_bench_sweep_sets.restype = None
        for name in ("aeo_source_iteration", "aeo_power_iteration", "aeo_time_march"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Tables:
    """Per-material cross-section tables + per-cell material index (see ae_oracle.c header)."""

    def __init__(self, mat, total, scat, chi, nusf, invspeed):
        self.mat = np.ascontiguousarray(mat, dtype=np.int32)
        self.total = _f64(total)
        self.chi = _f64(chi)
        self.nusf = _f64(nusf)
        self.invspeed = _f64(invspeed)
        self.M, self.G = self.total.shape
        self.scat = _f64(scat).reshape(self.M, self.G, self.G)
        self.Z = self.mat.shape[0]

    @classmethod
    def from_cell_arrays(cls, total_xs, scatter_xs, chi, nu_sigmaf, inv_speed):
        """One-group per-cell (Z,1) arrays (group.py:12-16) -> unique materials + index."""
        cols = np.column_stack([np.ravel(x) for x in (total_xs, scatter_xs, chi, nu_sigmaf, inv_speed)])
        uniq, inv = np.unique(cols, axis=0, return_inverse=True)
        return cls(inv.reshape(-1), uniq[:, 0:1], uniq[:, 1:2].reshape(-1, 1, 1), uniq[:, 2:3],
                   uniq[:, 3:4], uniq[:, 4:5])


def sweep(mu, hx, source, sigT):
    """group.py:41-60 for one angle.  source, sigT: (Z,) -> psi (Z,)."""
    source, sigT = _f64(np.ravel(source)), _f64(np.ravel(sigT))
    Z = source.shape[0]
    psi = np.empty(Z)
    lib().aeo_sweep(ctypes.c_int64(Z), ctypes.c_double(mu), ctypes.c_double(1 / hx),
                    _d(source), ctypes.c_int64(1), _d(sigT), ctypes.c_int64(1), _d(psi), ctypes.c_int64(1))
    return psi


_inflow_keep = None


def set_inflow(inflow):
    """Incident boundary fluxes (A, G) used by every sweep set from now on (OneDirection.boundary_condition,
    one_direction.py:12,24-31), or None for vacuum (what the reference always runs with)."""
    global _inflow_keep
    _inflow_keep = None if inflow is None else _f64(inflow)
    lib().aeo_set_inflow.restype = None
    lib().aeo_set_inflow(_d(_inflow_keep))


def set_reflect_left(on: bool):
    """Reflective LEFT boundary (symmetry plane at x = 0) for every sweep set from now on: a direction mu > 0 enters with
    the flux its mirror direction left with in the same sweep set.  False = the reference's vacuum."""
    lib().aeo_set_reflect_left.restype = None
    lib().aeo_set_reflect_left(1 if on else 0)


def sweep_set(mu, w, hx, q, sigT, want_psi=False):
    """All angles x groups with isotropic source q (Z,G); returns phi (Z,G) [, psi (Z,A,G)]."""
    q, sigT, mu, w = _f64(q), _f64(sigT), _f64(mu), _f64(w)
    Z, G = q.shape
    A = mu.shape[0]
    phi = np.empty((Z, G))
    psi = np.empty((Z, A, G)) if want_psi else None
    scratch = np.empty(2 * Z)
    lib().aeo_sweep_set(ctypes.c_int64(Z), A, G, _d(mu), _d(w), ctypes.c_double(hx), _d(q), _d(sigT),
                        _d(psi), _d(phi), _d(scratch))
    return (phi, psi) if want_psi else phi


def mg_sources(tab: Tables, phi):
    """(0.5 chi_g sum_g' nusf_g' phi_g', 0.5 sum_g' sigma_s(g'->g) phi_g') per cell, both (Z,G): the fission and
    scatter terms of the multigroup solvers (zone.py:38-51)."""
    phi = _f64(phi)
    Z, G = phi.shape
    fis, sca = np.empty((Z, G)), np.empty((Z, G))
    lib().aeo_mg_sources(ctypes.c_int64(Z), G, _i(tab.mat), _d(tab.scat), _d(tab.chi), _d(tab.nusf), _d(phi), _d(fis), _d(sca))
    return fis, sca


def source_iteration(mu, w, hx, tab: Tables, source, tol=1e-8, max_iterations=100, want_psi=False):
    """group.py:62-81.  Returns (status, phi_new (Z,G), iterations[, psi])."""
    mu, w, source = _f64(mu), _f64(w), _f64(source)
    Z, G, A = tab.Z, tab.G, mu.shape[0]
    phi = np.empty((Z, G))
    psi = np.empty((Z, A, G)) if want_psi else None
    it = ctypes.c_int(0)
    st = lib().aeo_source_iteration(ctypes.c_int64(Z), A, G, tab.M, _d(mu), _d(w), ctypes.c_double(hx),
                                    _i(tab.mat), _d(tab.total), _d(tab.scat), _d(source),
                                    ctypes.c_double(tol), max_iterations, _d(phi), _d(psi), ctypes.byref(it))
    return (st, phi, it.value, psi) if want_psi else (st, phi, it.value)


def power_iteration(mu, w, hx, tab: Tables, phi_old, tol=1e-8, max_iterations=100,
                    si_tol=1e-8, si_max_iterations=100, want_psi=False):
    """group.py:97-116.  Returns dict(status, k_new, k_old, iterations, phi_new, phi_old, k_hist, si_hist[, psi])."""
    mu, w = _f64(mu), _f64(w)
    Z, G, A = tab.Z, tab.G, mu.shape[0]
    phi_old = _f64(phi_old).reshape(Z, G).copy()
    phi_new = np.empty((Z, G))
    psi = np.empty((Z, A, G)) if want_psi else None
    k_new, k_old, it = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int(0)
    k_hist = np.zeros(max_iterations)
    si_hist = np.zeros(max_iterations, dtype=np.int32)
    st = lib().aeo_power_iteration(ctypes.c_int64(Z), A, G, tab.M, _d(mu), _d(w), ctypes.c_double(hx),
                                   _i(tab.mat), _d(tab.total), _d(tab.scat), _d(tab.chi), _d(tab.nusf),
                                   ctypes.c_double(tol), max_iterations, ctypes.c_double(si_tol),
                                   si_max_iterations, _d(phi_old), _d(phi_new), _d(psi),
                                   ctypes.byref(k_new), ctypes.byref(k_old), ctypes.byref(it),
                                   _d(k_hist), _i(si_hist))
    out = dict(status=st, k_new=k_new.value, k_old=k_old.value, iterations=it.value, phi_new=phi_new,
               phi_old=phi_old, k_hist=k_hist[:it.value], si_hist=si_hist[:it.value])
    if want_psi:
        out["psi"] = psi
    return out


def time_march(mu, w, hx, tab: Tables, S_ext, psi0, dt, num_steps, tol=1e-10,
               want_psi_snap=True, want_phi_snap=True):
    """time_unit.py:17-59 with repairs R2-R5.  psi0: (Z,A,G).  Returns dict."""
    mu, w = _f64(mu), _f64(w)
    Z, G, A = tab.Z, tab.G, mu.shape[0]
    S_ext = _f64(S_ext).reshape(Z, G)
    psi = _f64(psi0).reshape(Z, A, G).copy()
    phi_last = np.zeros((Z, G))
    psi_snap = np.zeros((Z, A, G, num_steps)) if want_psi_snap else None
    phi_snap = np.zeros((Z, G, num_steps)) if want_phi_snap else None
    outer = np.zeros(num_steps, dtype=np.int32)
    inner = np.zeros(num_steps, dtype=np.int32)
    st = lib().aeo_time_march(ctypes.c_int64(Z), A, G, tab.M, _d(mu), _d(w), ctypes.c_double(hx),
                              _i(tab.mat), _d(tab.total), _d(tab.scat), _d(tab.chi), _d(tab.nusf),
                              _d(tab.invspeed), _d(S_ext), ctypes.c_double(dt), num_steps,
                              ctypes.c_double(tol), _d(psi), _d(phi_last), _d(psi_snap), _d(phi_snap),
                              _i(outer), _i(inner))
    return dict(status=st, psi_last=psi, phi_last=phi_last, psi_snap=psi_snap, phi_snap=phi_snap,
                outer=outer, inner=inner)


def bench_sweep_sets(mu, w, hx, q, sigT, cdt, psi_prev, reps=1):
    """Throughput leg: `reps` time-dependent sweep sets on OMP_NUM_THREADS host threads."""
    mu, w, q, sigT, cdt, psi_prev = map(_f64, (mu, w, q, sigT, cdt, psi_prev))
    Z, G = q.shape
    A = mu.shape[0]
    psi_out = np.empty((Z, A, G))
    phi = np.empty((Z, G))
    lib().aeo_bench_sweep_sets(ctypes.c_int64(Z), A, G, _d(mu), _d(w), ctypes.c_double(hx), _d(q), _d(sigT),
                               _d(cdt), _d(psi_prev), _d(psi_out), _d(phi), reps)
    return phi, psi_out


# ------------------------------------------------------------------------------------------
# DMD: restatement of TimeUnit.compute_alpha_eigenvalues (time_unit.py:61-92).
# ------------------------------------------------------------------------------------------
def dmd_window(num_steps: int):
    """Column bookkeeping of time_unit.py:62-64,72,87.

    Returns (first, K): X = snapshots[first : first+K], Y = snapshots[first+1 : first+K+1] in
    ABSOLUTE step indices.  The reference applies `skip` twice (once when slicing psi, :65, and
    again when slicing reshaped_psi, :72), hence first = 2*skip.
    """
    skip = int(num_steps / 5)
    num_included = int(num_steps - skip - num_steps / 50)
    it = num_included - 1
    K = it - skip
    return 2 * skip, K


def dmd_alpha(snapshots, num_steps: int, dt: float, return_rank_info: bool = False):
    """snapshots: (M, num_steps) matrix, column s = flattened state of step s (any row order).

    Follows time_unit.py:61-91 operation for operation, INCLUDING the memory layout of the
    intermediate `reshaped_psi` (a C-contiguous (M, num_included) copy that is then sliced),
    because LAPACK/BLAS results depend on it at the 1e-16 level and the reference's rank rule
    (keep sigma_j while the tail sum exceeds 1e-13 of the total, :75-76) amplifies that to
    ~1e-7 in alpha on real marches ("DMD conditioning").
    """
    skip = int(num_steps / 5)                                   # :62
    num_included = int(num_steps - skip - num_steps / 50)       # :63
    it = num_included - 1                                       # :64
    M = snapshots.shape[0]
    reshaped = np.zeros((M, num_included))                      # :68
    for step in range(num_included):                            # :69-70
        reshaped[:, step] = snapshots[:, skip + step]
    U, S, V = np.linalg.svd(reshaped[:, skip:it], full_matrices=False)      # :72
    importance = 1 - np.cumsum(S) / np.sum(S)                   # :75
    S_reduced = S[importance > 1e-13]                           # :76
    r = len(S_reduced)
    S_square = np.zeros((r, r))
    S_square[np.diag_indices(r)] = 1 / S_reduced                # :77-80
    U = U[:, 0:r]
    V = V[0:r, :]
    psi_star = np.dot(U.conj().T, reshaped[:, skip + 1:it + 1])  # :87
    Atilde = np.dot(np.dot(psi_star, V.conj().T), S_square)     # :88-89
    eigs = np.linalg.eig(Atilde)[0]                             # :90
    alphas = (1 - 1.0 / eigs) / dt                              # :91
    if return_rank_info:
        return alphas, S, r
    return alphas


def dmd_alpha_from_psi(psi, num_steps: int, dt: float):
    """psi: (Z, A, G, steps) as TimeUnit.psi (time_unit.py:14); rows flattened C-order (:66-70)."""
    M = int(np.prod(psi.shape[:-1]))
    return dmd_alpha(psi.reshape(M, psi.shape[-1]), num_steps, dt)
