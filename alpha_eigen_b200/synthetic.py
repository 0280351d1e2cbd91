"""Synthetic multigroup slabs of the benchmark configurations.

The reference ships no runnable multigroup problem; these tables follow the recipe fixed in the
survey so that oracle, tests and bench all see the same inputs: three materials on the
Kornreich-Parsons five-region layout (slab.py:33-56 fractions 1/9, 2/9, 7/9, 8/9), seeded with
RandomState(2021) like the reference's own initial guess (phi.py:2).
"""
import numpy as np

from .material import TableMaterial


def synthetic_materials(G, seed=2021):
    """[fuel, moderator, absorber] TableMaterials.

    sigma_t,g ~ U(0.5,2)*(1+g/G); scatter matrix down-scatter dominant (plus in-group and a little
    up-scatter) with row sums c*sigma_t,g, c = 0.8 (fuel, moderator) / 0.1 (absorber) as
    material.py:12-20; nu_sigma_f,g = 0.7/G*U(0.5,1.5) in fuel only; chi normalised and
    concentrated in the 10 fastest groups; inv_speed log-spaced 1 -> 1e3.
    """
    rs = np.random.RandomState(seed)
    g = np.arange(G)
    d = g[None, :] - g[:, None]                             # to - from
    shape = np.where(d >= 0, np.exp(-d / 3.0), np.where(d >= -2, 1e-3, 0.0))
    chi = np.where(g < min(10, G), np.exp(-g / 3.0), 0.0)
    chi /= chi.sum()
    inv_speed = np.logspace(0, 3, G) if G > 1 else np.ones(1)
    out = []
    for name, c in (("fuel", 0.8), ("moderator", 0.8), ("absorber", 0.1)):
        total = rs.uniform(0.5, 2.0, G) * (1 + g / G)
        sc = shape * rs.uniform(0.5, 1.5, (G, G))
        sc *= (c * total / sc.sum(axis=1))[:, None]
        nusf = 0.7 / G * rs.uniform(0.5, 1.5, G) if name == "fuel" else np.zeros(G)
        out.append(TableMaterial(name, total, sc, nusf, chi, inv_speed))
    return out


def region_index(Z):
    """Material index per cell (0 fuel, 1 moderator, 2 absorber) on the scaled K-P layout."""
    x = (np.arange(Z) + 0.5) * (9.0 / Z)
    idx = np.full(Z, 2, dtype=np.int64)
    idx[(x <= 1) | (x >= 8)] = 0
    idx[((x > 1) & (x < 2)) | ((x > 7) & (x < 8))] = 1
    return idx


def synthetic_problem(Z, A, G, seed=2021, length=9.0):
    """Plain-array description usable by SlabDevice and by the oracle's Tables."""
    mats = synthetic_materials(G, seed)
    mu, w = np.polynomial.legendre.leggauss(A)
    stack = lambda name: np.array([getattr(m, name) for m in mats])
    return dict(Z=Z, A=A, G=G, hx=length / Z, mu=mu, weight=w, mat=region_index(Z), total=stack("total_xs"),
                scat=stack("scatter_xs"), chi=stack("chi"), nusf=stack("nu_sigmaf"), invspeed=stack("inv_speed"))
