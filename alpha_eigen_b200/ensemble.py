"""Ensembles of parameter-varied slabs (BASELINE.json configs[4]: 1024 twelve-group slabs with DMD).

Every member is an independent problem with its own convergence history, exactly as if it were run
alone, so results are identical to stand-alone solves.  Two ways to spread the work:

* member-parallel (default): members are dealt round-robin over the ranks (no communication at all) and,
  inside a rank, several members run concurrently on their own CUDA streams -- one small slab cannot fill
  a B200 (its sweep set is a few dozen CTAs), a dozen in flight can;
* row-sharded (`row_sharded=True`): every member is solved by all ranks together, angles dealt over the
  ranks, the tall psi snapshot matrix sharded by rows and the DMD factor merged across ranks
  (ae_tsqr_merge) -- the layout BASELINE.json names.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import distributed
from .device import SlabDevice, shard_angles
from .synthetic import synthetic_materials, region_index
from .time_unit import dmd_window


def synthetic_ensemble(E, Z, A, G, seed=2021, length=9.0):
    """Member j of the SURVEY section 8(d) ensemble: fuel nu_sigma_f scaled by 1 + 0.2 (j/(E-1) - 1/2),
    absorber scattering scaled by U(0.8, 1.2) drawn from RandomState(seed + j)."""
    mu, w = np.polynomial.legendre.leggauss(A)
    mat = region_index(Z)
    base = synthetic_materials(G, seed)
    stack = lambda name: np.array([getattr(m, name) for m in base])
    out = []
    for j in range(E):
        rs = np.random.RandomState(seed + j)
        nusf = stack("nu_sigmaf").copy()
        scat = stack("scatter_xs").copy()
        nusf[0] *= 1 + 0.2 * ((j / (E - 1) if E > 1 else 0.5) - 0.5)
        scat[2] *= rs.uniform(0.8, 1.2)
        half = rs.rand(Z, G)
        out.append(dict(Z=Z, A=A, G=G, hx=length / Z, mu=mu, weight=w, mat=mat, total=stack("total_xs"), scat=scat,
                        chi=stack("chi"), nusf=nusf, invspeed=stack("inv_speed"), phi0=half + half[::-1]))
    return out


def solve_member(p, dt, num_steps, comm=None, tol=1e-10, rank_tol=1e-13, keep_snapshots=False):
    """Time march + DMD of one member on the current CUDA stream.  Returns a dict with alphas etc."""
    import torch
    angles = shard_angles(p["mu"], comm.rank, comm.world) if comm is not None else None
    dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                     p["invspeed"], dt=dt, s_ext=p.get("s_ext"), angles=angles, comm=comm)
    psi0 = np.repeat(0.5 * p["phi0"][:, None, :], p["A"], axis=1)          # isotropic psi0 = phi0 / 2
    psi = dev.psi_to_device(psi0)
    n_psi = dev.G * dev.A * dev.Zp
    snap = torch.zeros(num_steps * n_psi, dtype=torch.float64, device=dev.dev)
    st, outer, inner = dev.time_march(psi, num_steps, tol, snap, None)
    out = dict(status=st, outer=outer, inner=inner)
    if st == 0:
        first, K = dmd_window(num_steps)
        R = dev.tsqr(snap, n_psi, first, K + 1)
        alphas, sigma, r = dev.dmd_from_r(R, K, dt, rank_tol)
        order = np.lexsort((alphas.imag, -alphas.real))
        out.update(alphas=alphas[order], sigma=sigma, rank=r)
    if keep_snapshots:
        out["psi_last"] = dev.psi_to_host(psi)
        out["angles"] = dev.angles
        out["snap"] = snap
        out["dev"] = dev
    return out


class Ensemble:
    def __init__(self, members, dt=0.1, num_steps=200, concurrent=8, row_sharded=False):
        self.members, self.dt, self.num_steps = members, dt, num_steps
        self.concurrent, self.row_sharded = max(int(concurrent), 1), row_sharded

    def run(self, keep_snapshots=False):
        """Solve every member; returns a list (one dict per member, same on every rank)."""
        import torch
        comm = distributed.communicator()
        world = comm.world if comm is not None else 1
        rank = comm.rank if comm is not None else 0
        E = len(self.members)
        if self.row_sharded and comm is not None:
            return [solve_member(p, self.dt, self.num_steps, comm=comm, keep_snapshots=keep_snapshots) for p in self.members]
        mine = list(range(rank, E, world))
        results = {}

        def work(j):
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                r = solve_member(self.members[j], self.dt, self.num_steps, keep_snapshots=keep_snapshots)
                stream.synchronize()
            return j, r

        with ThreadPoolExecutor(max_workers=self.concurrent) as pool:
            for j, r in pool.map(work, mine):
                results[j] = r
        if world > 1:
            import torch.distributed as dist
            slim = {j: {k: v for k, v in r.items() if k not in ("snap", "dev")} for j, r in results.items()}
            gathered = [None] * world
            dist.all_gather_object(gathered, slim)
            results = {j: r for part in gathered for j, r in part.items()}
        return [results[j] for j in range(E)]
