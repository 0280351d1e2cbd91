"""Ensembles of parameter-varied slabs (BASELINE.json configs[4]: 1024 twelve-group slabs with DMD).

Every member is an independent problem with its own convergence history, exactly as if it were run
alone, so results are identical to stand-alone solves.  Two ways to spread the work:

* member-parallel (default): members are dealt round-robin over the ranks (no communication at all).  Inside a
  rank the marches of all its members run in ONE cooperative kernel (`mode="batched"`, ae_ensemble_time_march):
  the grid is cut into teams of CTAs, one team per member in flight, each team pulling its next member from a
  device-side queue -- one small slab cannot fill a B200 (its sweep set is a few dozen CTAs), a dozen marching
  side by side with team-local barriers can.  `mode="streams"` is the earlier scheme (one fused kernel per
  member, several members in flight on their own CUDA streams), kept for comparison;
* row-sharded (`row_sharded=True`): every member is solved by all ranks together, angles dealt over the
  ranks, the tall psi snapshot matrix sharded by rows and the DMD factor merged across ranks
  (ae_tsqr_merge) -- the layout BASELINE.json names.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, distributed
from .device import SlabDevice
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
    dev = SlabDevice.sharded(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                             p["invspeed"], comm=comm, dt=dt, s_ext=p.get("s_ext"))
    psi0 = np.repeat(0.5 * p["phi0"][:, None, :], p["A"], axis=1)          # isotropic psi0 = phi0 / 2
    psi = dev.psi_to_device(psi0)
    n_psi = dev.Gl * dev.A * dev.Zp
    snap = torch.zeros(num_steps * n_psi, dtype=torch.float64, device=dev.dev)
    st, outer, inner = dev.time_march(psi, num_steps, tol, snap, None)
    out = dict(status=st, outer=outer, inner=inner)
    if st == 0:
        alphas, sigma, r = member_dmd(dev, snap, n_psi, num_steps, dt, rank_tol)    # row-sharded: factors merged in dev.tsqr
        out.update(alphas=alphas, sigma=sigma, rank=r)
    if keep_snapshots:
        out["psi_last"] = dev.psi_to_host(psi)
        out["angles"] = dev.angles
        out["groups"] = dev.groups
        out["snap"] = snap
        out["dev"] = dev
    return out


def member_dmd(dev, snap, row_len, num_steps, dt, rank_tol=1e-13):
    """DMD of one member's snapshot matrix [steps][row_len] (time_unit.py:61-92): (alphas sorted, sigma, rank)."""
    first, K = dmd_window(num_steps)
    R = dev.tsqr(snap, row_len, first, K + 1)
    alphas, sigma, r = dev.dmd_from_r(R, K, dt, rank_tol)
    order = np.lexsort((alphas.imag, -alphas.real))
    return alphas[order], sigma, r


def march_batched(devs, psis, psi_snaps, phi_snaps, num_steps, tol=1e-10):
    """ae_ensemble_time_march over same-shape members: devs [SlabDevice], psis [G][A][Zp] tensors (psi0 in, last psi
    out), psi_snaps / phi_snaps lists of tensors or None.  Returns (status [E], outer [E][steps], inner [E][steps],
    members in flight)."""
    import torch
    E = len(devs)
    d0 = devs[0]
    hist = torch.zeros(E, 2 * num_steps, dtype=torch.int32, device=d0.dev)
    arr = (_lib.AeMember * E)()
    ptr = lambda t: t.data_ptr() if t is not None else None
    for j, dev in enumerate(devs):
        arr[j] = _lib.AeMember(C.pointer(dev.slab), C.pointer(dev.work), psis[j].data_ptr(),
                               ptr(psi_snaps[j]) if psi_snaps is not None else None,
                               ptr(phi_snaps[j]) if phi_snaps is not None else None,
                               hist[j].data_ptr())
    status = (C.c_int32 * E)()
    teams = C.c_int32(0)
    _lib.check(d0.lib.ae_ensemble_time_march(arr, E, num_steps, tol, status, C.byref(teams), d0.stream))
    h = hist.cpu().numpy()
    return np.array(status[:], dtype=np.int32), h[:, :num_steps].copy(), h[:, num_steps:].copy(), teams.value


def solve_members_batched(members, dt, num_steps, tol=1e-10, rank_tol=1e-13, keep_snapshots=False, snapshots="psi",
                          memory_fraction=0.6, timings=None):
    """March a list of same-shape members with the batched kernel (in chunks that fit HBM), then their DMDs.
    `timings` (dict, optional) receives wall seconds of the setup / march / dmd phases."""
    import time
    import torch
    results = []
    clock = {"setup": 0.0, "march": 0.0, "dmd": 0.0}

    def lap(name, t0):
        torch.cuda.synchronize()
        clock[name] += time.perf_counter() - t0
        return time.perf_counter()
    p0 = members[0]
    free, _ = torch.cuda.mem_get_info()
    Zp = int(_lib.load().ae_padded_cells(p0["Z"]))
    rows = p0["G"] * Zp * (p0["A"] if snapshots == "psi" else 1)
    per_member = 8 * (rows * num_steps + p0["G"] * Zp * (p0["A"] + 12))
    chunk = max(1, min(len(members), int(memory_fraction * free / per_member)))
    for c0 in range(0, len(members), chunk):
        part = members[c0:c0 + chunk]
        devs, psis, snaps = [], [], []
        t0 = time.perf_counter()
        for p in part:
            dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                             p["invspeed"], dt=dt, s_ext=p.get("s_ext"))
            psi0 = np.repeat(0.5 * p["phi0"][:, None, :], p["A"], axis=1)      # isotropic psi0 = phi0 / 2
            devs.append(dev)
            psis.append(dev.psi_to_device(psi0))
            snaps.append(torch.zeros(num_steps * rows, dtype=torch.float64, device=dev.dev))
        t0 = lap("setup", t0)
        status, outer, inner, teams = march_batched(devs, psis, snaps if snapshots == "psi" else None,
                                                    snaps if snapshots == "phi" else None, num_steps, tol)
        t0 = lap("march", t0)
        # DMDs: one TSQR per member (each fills the GPU), then the small dense parts of ALL members in one launch
        # (ae_dmd_from_r_batched: one thread block per window)
        first, K = dmd_window(num_steps)
        ok = [j for j in range(len(devs)) if status[j] == 0]
        dmds = {}
        if ok:
            d0 = devs[0]
            r_stack = torch.empty(len(ok) * (K + 1) * (K + 1), dtype=torch.float64, device=d0.dev)
            for i, j in enumerate(ok):
                r_stack[i * (K + 1) * (K + 1):(i + 1) * (K + 1) * (K + 1)] = devs[j].tsqr(snaps[j], rows, first, K + 1)
            for j, res in zip(ok, d0.dmd_from_r_batched(r_stack, len(ok), K, dt, rank_tol)):
                if isinstance(res, int):
                    raise RuntimeError(f"DMD of member {c0 + j} failed with status {res}")
                alphas, sigma, r = res
                order = np.lexsort((alphas.imag, -alphas.real))
                dmds[j] = (alphas[order], sigma, r)
        for j, dev in enumerate(devs):
            out = dict(status=int(status[j]), outer=outer[j], inner=inner[j], in_flight=teams)
            if j in dmds:
                out.update(alphas=dmds[j][0], sigma=dmds[j][1], rank=dmds[j][2])
            if keep_snapshots:
                out.update(psi_last=dev.psi_to_host(psis[j]), angles=dev.angles, snap=snaps[j], dev=dev)
            results.append(out)
        lap("dmd", t0)
        del devs, psis, snaps
    if timings is not None:
        timings.update(clock)
    return results


class Ensemble:
    def __init__(self, members, dt=0.1, num_steps=200, concurrent=8, row_sharded=False, mode="batched", snapshots="psi"):
        self.members, self.dt, self.num_steps = members, dt, num_steps
        self.concurrent, self.row_sharded = max(int(concurrent), 1), row_sharded
        assert mode in ("batched", "streams") and snapshots in ("psi", "phi")
        self.mode, self.snapshots = mode, snapshots

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
        if self.mode == "batched" and mine:
            got = solve_members_batched([self.members[j] for j in mine], self.dt, self.num_steps,
                                        keep_snapshots=keep_snapshots, snapshots=self.snapshots)
            results = dict(zip(mine, got))
            mine = []

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
