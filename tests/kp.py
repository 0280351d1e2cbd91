"""Kornreich-Parsons 9-mfp slab as plain arrays, for oracle-level tests.

Restates slab.py:6-12,33-56 (geometry, Gauss-Legendre quadrature, 5-region material map),
material.py:9-20 (one-group cross-sections) and phi.py:2,10-11 (seeded symmetric initial guess)
without importing the reference or the product package.
"""
import numpy as np

from oracle.ae_oracle import Tables

XS = {  # material.py:9-20: (total, scatter, chi, nu_sigmaf, inv_speed)
    "fuel": (1.0, 0.8, 1.0, 0.7, 1.0),
    "moderator": (1.0, 0.8, 1.0, 0.0, 1.0),
    "absorber": (1.0, 0.1, 1.0, 0.0, 1.0),
}


def kp_region(mid):
    """slab.py:41-56 with the edges of slab.py:33-39 (1, 2, 7, 8, 9 mfp)."""
    if mid <= 1:
        return "fuel"
    if 1 < mid < 2:
        return "moderator"
    if 2 <= mid <= 7:
        return "absorber"
    if 7 < mid < 8:
        return "moderator"
    if 8 <= mid <= 9:
        return "fuel"
    raise ValueError(mid)


def kp_problem(length=9, cells_per_mfp=50, num_angles=4, nusf_fuel=0.7):
    Z = length * cells_per_mfp                                  # slab.py:8
    hx = length / Z                                             # slab.py:9
    mu, weight = np.polynomial.legendre.leggauss(num_angles)    # slab.py:11
    mids = np.linspace(hx / 2, length - hx / 2, Z)              # group.py:25
    names = [kp_region(m) for m in mids]
    cols = np.array([XS[n] for n in names])
    cols[:, 3] = np.where(cols[:, 3] > 0, nusf_fuel, 0.0)
    total, scatter, chi, nusf, inv = (cols[:, j:j + 1].copy() for j in range(5))
    phi = np.random.RandomState(2021).rand(Z, 1)                # phi.py:2,10 in a fresh process
    phi0 = phi + np.flip(phi, 0)                                # phi.py:11
    return dict(Z=Z, A=num_angles, G=1, hx=hx, mu=mu, weight=weight, total_xs=total, scatter_xs=scatter,
                chi=chi, nu_sigmaf=nusf, inv_speed=inv, phi0=phi0, regions=names,
                tables=Tables.from_cell_arrays(total, scatter, chi, nusf, inv))


def kp_psi0(p):
    """Initial condition of the repaired time march: isotropic psi0 = phi0/2 (sum of weights is 2)."""
    return np.repeat(0.5 * p["phi0"][:, None, :], p["A"], axis=1)
