"""The block kernel's numerical code (alpha_eigen_b200/csrc/ae_dmd_small.cuh: one-sided Jacobi SVD, rank rule, Atilde,
balance + Hessenberg + double-shift QR) instantiated with a one-thread team on the host (tests/dmd_small_host.cpp) and
checked against LAPACK and the oracle's DMD: the GPU runs the same templates with a thread block as the team."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import ae_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dmd_small") / "dmd_small_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "dmd_small_host.cpp")], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _match(ref, got):
    """largest relative distance after nearest-neighbour pairing"""
    got, err = list(got), 0.0
    for x in ref:
        k = int(np.argmin([abs(x - y) for y in got]))
        err = max(err, abs(x - got[k]) / max(abs(x), 1e-300))
        got.pop(k)
    return err


def _eig(host, a):
    n = a.shape[0]
    a = np.ascontiguousarray(a, dtype=np.float64)
    wr, wi = np.zeros(n), np.zeros(n)
    assert host.eig_small_host(_p(a), n, _p(wr), _p(wi)) == 0
    return wr + 1j * wi


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 17, 33, 64, 115])
def test_eigenvalues_match_lapack(host, n):
    a = np.random.RandomState(n).randn(n, n)
    assert _match(np.linalg.eigvals(a), _eig(host, a)) < 1e-11          # random matrices: eigenvalue condition numbers ~ n


def test_eigenvalues_of_a_badly_scaled_matrix(host):
    """Atilde = U^T Y V S^-1 is graded like sigma_i / sigma_j: without balancing the small eigenvalues drown."""
    rs = np.random.RandomState(0)
    a = rs.randn(30, 30)
    d = 10.0 ** rs.uniform(-8, 8, 30)
    b = (a * d[None, :]) / d[:, None]                                   # similar to a
    assert _match(np.linalg.eigvals(a), _eig(host, b)) < 1e-12


def test_companion_matrix_complex_pairs(host):
    roots = np.array([0.9, 0.5 + 0.3j, 0.5 - 0.3j, -0.2 + 0.7j, -0.2 - 0.7j, 0.1])
    a = np.zeros((6, 6))
    a[0, :] = -np.poly(roots)[1:].real
    a[np.arange(1, 6), np.arange(0, 5)] = 1.0
    assert _match(roots, _eig(host, a)) < 1e-12


def _window(host, R, K, dt, tol=1e-13):
    R = np.ascontiguousarray(R)
    are, aim, sig, rk = np.zeros(K), np.zeros(K), np.zeros(K), C.c_int(0)
    st = host.dmd_small_host(_p(R), K, C.c_double(dt), C.c_double(tol), _p(are), _p(aim), _p(sig), C.byref(rk))
    return st, are[:rk.value] + 1j * aim[:rk.value], sig, rk.value


def test_dmd_windows_of_the_fixtures(host, golden_dir):
    from alpha_eigen_b200.time_unit import dmd_window
    seen = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, "dmd_*.npz"))):
        g = np.load(f)
        if "psi" not in g.files:
            continue
        seen += 1
        psi, steps, dt = g["psi"], int(g["steps"]), float(g["dt"])
        first, K = dmd_window(steps)
        W = psi.reshape(-1, steps)[:, first:first + K + 1]
        Rr = np.linalg.qr(W, mode="r")
        R = np.zeros((K + 1, K + 1))
        R[:Rr.shape[0]] = Rr
        st, al, sig, rk = _window(host, R, K, dt)
        ref = g["alphas"]
        assert st == 0 and rk == len(ref)                               # same rank decision (time_unit.py:75-76)
        sv = np.linalg.svd(W[:, :K], compute_uv=False)
        m = min(len(sv), K)
        assert np.max(np.abs(sig[:m] - sv[:m])) < 2e-15 * sv[0]
        if "march" in os.path.basename(f):
            # the march window is ill conditioned: the reference's own dominant alpha moves by ~1e-5 under a
            # 1e-15 relative perturbation of the snapshots (test_dmd_march_fixture_conditioning_aware measures it)
            srt = lambda z: z[np.lexsort((z.imag, -z.real))]
            assert abs(srt(al)[0] - srt(ref)[0]) < 1e-4 * abs(srt(ref)[0])
        else:
            assert _match(ref, al) < 1e-9
    assert seen >= 3


def test_no_singular_value_survives(host):
    st, _, _, rk = _window(host, np.zeros((5, 5)), 4, 0.1)
    assert st == 3 and rk == 0                                          # kNoRank


def _reference_window(W, K, dt, tol):
    """time_unit.py:72-91 on a window W (M, K+1) with LAPACK: the arithmetic the block kernel replaces."""
    U, S, Vt = np.linalg.svd(W[:, :K], full_matrices=False)
    r = int(np.sum(1 - np.cumsum(S) / np.sum(S) > tol))
    At = U[:, :r].T @ W[:, 1:K + 1] @ Vt[:r].T @ np.diag(1 / S[:r])
    return (1 - 1 / np.linalg.eigvals(At)) / dt, r


@pytest.mark.parametrize("K", [9, 28, 57, 115])
def test_known_linear_dynamics(host, K):
    """Snapshots of a linear system with five strong modes (one complex pair) and a weak sixth: the reference's rule keeps
    five (it always drops the last mode before an exactly zero tail), and their alphas are (1 - 1/lambda)/dt
    (time_unit.py:91) up to the weak mode's leakage; against LAPACK on the same window the spectrum agrees to rounding."""
    rs = np.random.RandomState(K)
    lam = np.array([0.999, 0.97, 0.9 + 0.2j, 0.9 - 0.2j, 0.8])
    M, dt = 400, 0.1
    v = rs.randn(M, 4)
    v_re, v_im = rs.randn(M), rs.randn(M)
    amp = rs.rand(4) + 0.5
    W = np.zeros((M, K + 1))
    for s in range(K + 1):
        pair = (lam[2] ** s) * (v_re + 1j * v_im)
        W[:, s] = (amp[0] * 0.999 ** s * v[:, 0] + amp[1] * 0.97 ** s * v[:, 1] + 2 * amp[2] * pair.real
                   + amp[3] * 0.8 ** s * v[:, 2] + 1e-7 * 0.5 ** s * v[:, 3])
    Rr = np.linalg.qr(W, mode="r")
    R = np.zeros((K + 1, K + 1))
    R[:Rr.shape[0]] = Rr
    st, al, sig, rk = _window(host, R, K, dt, tol=1e-10)
    ref, r_ref = _reference_window(W, K, dt, 1e-10)
    assert st == 0 and rk == r_ref and rk in (4, 5)                      # the 9-column window is too short for the fifth
    assert _match(ref, al) < 1e-8
    if rk == 5:
        assert _match((1 - 1 / lam) / dt, al) < 1e-4


def test_rank_deficient_wide_window(host):
    """More columns than rows (K > M, as in the 40-cell fixture): exact zero singular values, rank from the rule."""
    rs = np.random.RandomState(5)
    M, K = 30, 60
    W = rs.randn(M, 3) @ rs.randn(3, K + 1)
    Rr = np.linalg.qr(W, mode="r")
    R = np.zeros((K + 1, K + 1))
    R[:Rr.shape[0]] = Rr
    st, al, sig, rk = _window(host, R, K, 0.1, tol=1e-10)
    ref, r_ref = _reference_window(W, K, 0.1, 1e-10)
    assert st == 0 and rk == r_ref
    sv = np.linalg.svd(W[:, :K], compute_uv=False)
    assert np.max(np.abs(sig[:3] - sv[:3])) < 1e-14 * sv[0] and np.all(sig[3:] < 1e-13 * sv[0])
