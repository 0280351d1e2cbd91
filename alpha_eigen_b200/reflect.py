"""A reflective LEFT boundary (symmetry plane at x = 0) for the slab solvers, by the method of images.

The reference has no reflective option: `OneDirection.boundary_condition` is the only boundary datum and it is always 0
(one_direction.py:12,24-31; SURVEY section 8 f3 lists reflective boundaries as the next geometry feature).  A slab [0, L]
whose left face reflects is the right half of the slab [-L, L] with the materials, sources and initial fluxes mirrored and
vacuum on both faces: diamond differences on a mirrored mesh with a symmetric quadrature are exactly symmetric, so the
right half of that solve satisfies psi(0, mu) = psi(0, -mu) cell for cell, and the convergence norms of the doubled problem
are the half problem's (numerator and denominator both double), so the iteration counts are the same.

This costs twice the sweep work of an in-kernel reflection (the sweep kernels would have to publish a row's exiting edge
flux and start the mirror row from it) and changes no kernel: the doubled problem runs through the
ordinary `SlabDevice`.  The oracle implements the TRUE reflection (oracle/ae_oracle.c: aeo_set_reflect_left), which
reproduces the reference's golden k of the symmetric K-P slab from its right half alone; the tests compare this module
with it.
"""
from __future__ import annotations

import numpy as np

from .device import SlabDevice


def mirror_cells(a):
    """(Z, ...) on [0, L] -> (2Z, ...) on [-L, L]."""
    a = np.asarray(a)
    return np.concatenate([a[::-1], a], axis=0)


def mirror_angles(mu):
    """index of -mu[n] for every n (symmetric quadrature required)."""
    mu = np.asarray(mu, dtype=np.float64)
    idx = np.array([int(np.argmin(np.abs(mu + m))) for m in mu])
    if not np.all(np.abs(mu[idx] + mu) <= 1e-14 * np.abs(mu)):
        raise ValueError("reflective boundary needs a quadrature that is symmetric in mu")
    return idx


def mirror_psi(psi, mu):
    """(Z, A, G) on [0, L] -> (2Z, A, G) on [-L, L]: psi(-x, mu) = psi(x, -mu)."""
    psi = np.asarray(psi)
    return np.concatenate([psi[::-1][:, mirror_angles(mu), :], psi], axis=0)


class ReflectedSlabDevice:
    """The half slab [0, L] (Z cells) with a reflective left face and a vacuum right face.  Host arrays in, host arrays of
    the half slab out; constructor arguments as `SlabDevice` (mat, s_ext, ... given on the half slab)."""

    def __init__(self, Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, dt=None, s_ext=None, **kw):
        if kw.get("inflow") is not None:
            raise ValueError("incident fluxes of the doubled slab would enter through the image's face: not supported")
        self.Z, self.mu = int(Z), np.asarray(mu, dtype=np.float64)
        mirror_angles(self.mu)
        self.full = SlabDevice(2 * self.Z, hx, mu, weight, mirror_cells(mat), total, scat, chi, nusf, invspeed, dt=dt,
                               s_ext=None if s_ext is None else mirror_cells(s_ext), **kw)

    def _half(self, a):
        return np.ascontiguousarray(a[self.Z:])

    def source_iteration(self, source, tol=1e-8, max_iterations=100):
        """group.py:62-81 with `source` (Z, G) as the fixed source.  Returns (status, phi (Z, G), iterations)."""
        f = self.full
        st, it = f.source_iteration(f.cells_to_device(mirror_cells(source)), tol=tol, max_iterations=max_iterations)
        return st, self._half(f.cells_to_host(f.phi)), it

    def power_iteration(self, phi_old, **kw):
        """group.py:97-116 from the half slab's phi.old (Z, G); dict like `SlabDevice.power_iteration`, half-slab fluxes."""
        r = self.full.power_iteration(mirror_cells(phi_old), **kw)
        r["phi_new"], r["phi_old"] = self._half(r["phi_new"]), self._half(r["phi_old"])
        return r

    def time_march(self, psi0, num_steps, tol=1e-10):
        """time_unit.py:17-27 from psi0 (Z, A, G).  Returns dict(status, outer, inner, psi_last (Z, A, G), phi_last (Z, G))."""
        f = self.full
        psi = f.psi_to_device(mirror_psi(psi0, self.mu))
        st, outer, inner = f.time_march(psi, num_steps, tol)
        got = f.psi_to_host(psi)                                        # device angle order
        last = np.empty_like(got)
        last[:, f.angles, :] = got                                      # back to the quadrature's order
        return dict(status=st, outer=outer, inner=inner, psi_last=self._half(last),
                    phi_last=self._half(f.cells_to_host(f.phi)))
