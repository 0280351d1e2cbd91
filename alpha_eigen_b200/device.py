"""Device-resident slab problem: torch owns HBM and streams, the C ABI does the arithmetic.

`SlabDevice` is the single place where the Python classes that mirror the reference
(group.py / time_unit.py) meet libalpha_eigen_b200.so.  It uploads the cross-section tables
and per-cell arrays in the lane-blocked storage order once, allocates the work buffers, and
exposes one method per C entry point.  There is no CPU path: every method needs CUDA.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, layout


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.AlphaEigenError("alpha_eigen_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def shard_angles(mu: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Reference-angle indices owned by `rank`: negatives first, dealt round-robin so every rank
    gets the same mix of directions (the rows of one sweep set are independent, group.py:74-77)."""
    order = np.concatenate([np.flatnonzero(mu < 0), np.flatnonzero(mu >= 0)])
    mine = order[rank::world]
    return np.concatenate([mine[mu[mine] < 0], mine[mu[mine] >= 0]])


def group_range(G: int, rank: int, world: int):
    """(first group, number of groups) of `rank` in the balanced contiguous partition of G groups (ae_group_range)."""
    base, rem = divmod(int(G), int(world))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_mode_for(G: int, world: int) -> str:
    """How `world` ranks split the (angle x group) rows of a sweep set.  With at least two groups
    per rank (and more groups than the one-thread-per-cell update kernel handles) ranks own whole GROUPS with all
    angles: a group's flux needs no reduction and the only exchange is one all-gather of the new flux.  Otherwise
    (one-group problems, few groups) ranks own ANGLES of every group and the flux is reduce-scattered."""
    if world <= 1:
        return "single"
    return "group" if (G >= 2 * world and G > 8) else "angle"


class Comm:
    """NCCL communicator created from a torch.distributed rendezvous (one process per GPU)."""

    def __init__(self, rank: int, world: int):
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(lib.ae_comm_unique_id(ident))
        t = torch.tensor(list(bytes(ident)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.handle = C.c_void_p()
        _lib.check(lib.ae_comm_init(C.byref(self.handle), rank, world, ident))
        self.rank, self.world = rank, world

    def allreduce_(self, tensor):
        torch = _torch()
        _lib.check(_lib.load().ae_comm_allreduce_sum(self.handle, tensor.data_ptr(), tensor.numel(),
                                                    torch.cuda.current_stream().cuda_stream))
        return tensor

    def close(self):
        if self.handle:
            _lib.load().ae_comm_destroy(self.handle)
            self.handle = C.c_void_p()


class HostPipe:
    """Double-buffered, sharded host I/O for a stream of steps whose sources come from (and whose fluxes go to)
    pinned host memory: the upload of step i+1's source and the download of step i-1's flux run on their own CUDA
    streams (the copy engines) while step i's kernels run, so in steady state a step costs max(kernels, copies)
    instead of their sum.  A rank moves only ITS share: its cell range of every group when the sweeps are
    angle-sharded (the ranges are then all-gathered on the device), its own groups when they are group-sharded (no
    exchange at all: a rank only ever reads the source rows of the groups it sweeps).  Usage per step i:

        pipe.feed(i, rows)              # main stream: wait for upload i, pack it into rows [G][Zp] (+ all-gather)
        pipe.upload(i + 1, next_host)   # copy stream: H2D of the next source, overlaps the kernels below
        ... kernels of step i ...
        pipe.drain(i, flux_rows, out_host)   # main: unpack this rank's share; copy stream: D2H into out_host

    and pipe.wait() before the host touches the last results.  upload(0, ...) primes the pipe."""

    def __init__(self, dev: "SlabDevice"):
        torch = dev.torch
        self.dev = dev
        n = max(self.shard_shape()[0], 1) * self.shard_shape()[1]
        f64 = dict(dtype=torch.float64, device=dev.dev)
        self.stage_in = [torch.empty(n, **f64) for _ in range(2)]
        self.stage_out = [torch.empty(n, **f64) for _ in range(2)]
        self.copy_in, self.copy_out = torch.cuda.Stream(device=dev.dev), torch.cuda.Stream(device=dev.dev)
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]
        self.in_ready, self.in_free, self.out_ready, self.out_free = ev(), ev(), ev(), ev()
        self.in_used, self.out_used = [False, False], [False, False]

    def shard_shape(self):
        """(cells, groups) of the host arrays this rank uploads / downloads per step."""
        dev = self.dev
        if dev.gl > 0:
            return dev.Z, dev.gl
        return dev.cell_range()[1], dev.G

    def upload(self, i, shard_host):
        """Start the H2D copy of step i's source: pinned (cells, groups) tensor with this rank's share."""
        torch, k = self.dev.torch, i & 1
        with torch.cuda.stream(self.copy_in):
            if self.in_used[k]:
                self.copy_in.wait_event(self.in_free[k])            # step i-2 has packed this slot
            self.stage_in[k][:shard_host.numel()].copy_(shard_host.view(-1), non_blocking=True)
            self.in_ready[k].record(self.copy_in)
        self.in_used[k] = True

    def feed(self, i, rows):
        """On the current stream: pack the uploaded source of step i into `rows` [G][Zp]."""
        dev, k = self.dev, i & 1
        main = dev.torch.cuda.current_stream(dev.dev)
        main.wait_event(self.in_ready[k])
        if dev.gl > 0:                                              # own groups, all cells: rows [g0, g0+gl)
            _lib.check(dev.lib.ae_pack_cells(self.stage_in[k].data_ptr(), dev.Z, dev.Zp, dev.gl,
                                             rows.data_ptr() + 8 * dev.g0 * dev.Zp, dev.stream))
            self.in_free[k].record(main)
            return
        _, count, s0, cells = dev.cell_range()
        _lib.check(dev.lib.ae_pack_cells_range(self.stage_in[k].data_ptr(), count, cells, dev.Zp, dev.G,
                                               rows.data_ptr() + 8 * s0, dev.stream))
        self.in_free[k].record(main)
        if dev.comm is not None:
            _lib.check(dev.lib.ae_comm_gather_rows(dev.comm.handle, rows.data_ptr(), dev.G, dev.Zp, dev.stream))

    def drain(self, i, rows, shard_host):
        """Unpack this rank's share of `rows` [G][Zp] (current stream) and start its D2H copy."""
        dev, k = self.dev, i & 1
        torch = dev.torch
        main = torch.cuda.current_stream(dev.dev)
        if self.out_used[k]:
            main.wait_event(self.out_free[k])                       # the download of step i-2 has left this slot
        if dev.gl > 0:
            _lib.check(dev.lib.ae_unpack_cells(rows.data_ptr() + 8 * dev.g0 * dev.Zp, dev.Z, dev.Zp, dev.gl,
                                               self.stage_out[k].data_ptr(), dev.stream))
        else:
            _, count, s0, cells = dev.cell_range()
            _lib.check(dev.lib.ae_unpack_cells_range(rows.data_ptr() + 8 * s0, count, cells, dev.Zp, dev.G,
                                                     self.stage_out[k].data_ptr(), dev.stream))
        self.out_ready[k].record(main)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.out_ready[k])
            shard_host.view(-1).copy_(self.stage_out[k][:shard_host.numel()], non_blocking=True)
            self.out_free[k].record(self.copy_out)
        self.out_used[k] = True

    def join(self):
        """Make the current stream wait for every copy issued so far (for device-side timing)."""
        main = self.dev.torch.cuda.current_stream(self.dev.dev)
        main.wait_stream(self.copy_in)
        main.wait_stream(self.copy_out)

    def wait(self):
        self.copy_in.synchronize()
        self.copy_out.synchronize()


class SlabDevice:
    """One (slab, cross-sections, dt) problem resident in HBM.

    mu, weight ... all quadrature angles of the problem (slab.py:11); `angles` selects the
                   reference-angle indices this rank owns (default: all)
    mat .......... (Z,) material index per cell
    total/chi/nusf/invspeed (M,G), scat (M,G,G) with scat[m][g'][g] = sigma_s(g'->g)
    dt ........... None for steady state; otherwise sigT = total + inv_speed/dt (time_unit.py:24)
    s_ext ........ (Z,G) fixed external source or None
    inflow ....... (A_total, G) or (A_total,) flux entering the slab at each angle's upstream boundary
                   (OneDirection.boundary_condition, one_direction.py:12); None = vacuum
    """

    def __init__(self, Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, dt=None, s_ext=None,
                 angles=None, comm: Comm | None = None, device=None, batch=None, flags=0, group_range_=None, inflow=None):
        torch = _torch()
        self.torch = torch
        self.lib = _lib.load()
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        mu = np.asarray(mu, dtype=np.float64)
        weight = np.asarray(weight, dtype=np.float64)
        if np.any(np.abs(mu) <= 1e-10):
            raise AssertionError("abs(mu) must exceed 1e-10")          # group.py:44
        if angles is None:
            angles = shard_angles(mu, 0, 1)
        self.angles = np.asarray(angles, dtype=np.int64)              # device angle a <-> reference angle angles[a]
        self.Z, self.hx = int(Z), float(hx)
        world = comm.world if comm is not None else 1
        # group-sharded: this rank sweeps groups [g0, g0+gl) with all angles; cells are not sharded then
        self.g0, self.gl = (int(group_range_[0]), int(group_range_[1])) if group_range_ is not None else (0, 0)
        self.shard_mode = "single" if comm is None else ("group" if self.gl > 0 else "angle")
        self.Zp = layout.padded_cells(self.Z, 1 if self.gl > 0 else world)   # angle mode: equal 256-cell-aligned cell ranges per rank
        self.A_total = len(mu)
        self.A = len(self.angles)
        total = np.asarray(total, dtype=np.float64)
        self.M, self.G = total.shape
        if self.gl > 0:
            assert comm is not None and 0 <= self.g0 and self.g0 + self.gl <= self.G
        self.Gl = self.gl if self.gl > 0 else self.G                  # groups whose psi / partials live on this rank
        self.groups = np.arange(self.g0, self.g0 + self.Gl)
        self.dt = dt
        self.comm = comm
        mu_loc, w_loc = mu[self.angles], weight[self.angles]
        self.n_neg = int(np.sum(mu_loc < 0))
        assert np.all(mu_loc[:self.n_neg] < 0) and np.all(mu_loc[self.n_neg:] > 0)
        ihx = 1 / self.hx                                             # group.py:46
        mat = np.asarray(mat, dtype=np.int64)
        sigT = total[mat]                                             # (Z,G)
        inv_cell = np.asarray(invspeed, dtype=np.float64)[mat]
        if dt is not None:
            sigT = sigT + 1 / dt * inv_cell                           # time_unit.py:24
        up = self._upload
        self.t_mu = up(mu_loc * ihx)
        self.t_w = up(w_loc)
        self.t_aidx = up(self.angles.astype(np.int32))
        self.t_mat = up(layout.to_storage(mat.astype(np.int32), self.Z, self.Zp))
        self.t_scat = up(np.asarray(scat, dtype=np.float64).reshape(self.M, self.G, self.G))
        self.t_chi = up(np.asarray(chi, dtype=np.float64).reshape(self.M, self.G))
        self.t_nusf = up(np.asarray(nusf, dtype=np.float64).reshape(self.M, self.G))
        self.t_hs = self.cells_to_device(0.5 * sigT)                  # [G][Zp]
        self.t_cdt = self.cells_to_device(inv_cell / dt) if dt is not None else None
        self.t_sext = self.cells_to_device(s_ext) if s_ext is not None else None
        self.t_inflow = None
        if inflow is not None and np.any(np.asarray(inflow) != 0):
            bc = np.asarray(inflow, dtype=np.float64).reshape(self.A_total, -1)
            bc = np.broadcast_to(bc, (self.A_total, self.G))
            self.t_inflow = up(np.ascontiguousarray(bc[self.angles].T))          # [G][A] local angles
        del sigT, inv_cell
        self.P = int(self.lib.ae_partial_count(self.A, self.n_neg))
        n = self.G * self.Zp
        f64 = dict(dtype=torch.float64, device=self.dev)
        self._phi_bufs = [torch.zeros(n, **f64), torch.zeros(n, **f64)]      # ping-pong pair, see `phi`
        self.phi_outer = torch.zeros(n, **f64)
        self.reduced = torch.zeros(n, **f64) if (comm is not None and self.gl == 0) else None
        self.phi_fixed = torch.zeros(n, **f64) if dt is not None else None
        self.base = torch.zeros(n, **f64)
        self.q0 = torch.zeros(n, **f64)
        self.q1 = torch.zeros(n, **f64)
        self.partial = torch.zeros(self.P * self.Gl * self.Zp, **f64)
        self.scratch = torch.zeros(4 * int(self.lib.ae_norm_blocks(self.Zp)) + 2 * self.G, **f64)
        self.sync = torch.zeros(int(self.lib.ae_sweep_sync_bytes()), dtype=torch.uint8, device=self.dev)
        self.ctrl = torch.zeros(C.sizeof(_lib.AeCtrl), dtype=torch.uint8, device=self.dev)
        self.ctrl_host = torch.zeros(C.sizeof(_lib.AeCtrl), dtype=torch.uint8).pin_memory()
        if batch is None:       # small problems are launch-latency bound: poll the flag less often
            batch = 1 if (comm is not None or self.Z * self.A * self.G > (1 << 24)) else 8
        self.slab = _lib.AeSlab(self.Z, self.Zp, self.A, self.n_neg, self.G, self.M, self.t_mu.data_ptr(),
                                self.t_w.data_ptr(), self.t_mat.data_ptr(), self.t_scat.data_ptr(),
                                self.t_chi.data_ptr(), self.t_nusf.data_ptr(), self.t_hs.data_ptr(),
                                self.t_cdt.data_ptr() if self.t_cdt is not None else None,
                                self.t_sext.data_ptr() if self.t_sext is not None else None, None,
                                self.t_inflow.data_ptr() if self.t_inflow is not None else None, self.g0, self.gl)
        self._build_tile_info()
        self.work = _lib.AeWork(self._phi_bufs[0].data_ptr(), self._phi_bufs[1].data_ptr(), self.phi_outer.data_ptr(),
                                self.base.data_ptr(), self.q0.data_ptr(), self.q1.data_ptr(), self.partial.data_ptr(),
                                self.reduced.data_ptr() if self.reduced is not None else None,
                                self.phi_fixed.data_ptr() if self.phi_fixed is not None else None,
                                self.scratch.data_ptr(), self.ctrl.data_ptr(), self.ctrl_host.data_ptr(),
                                self.sync.data_ptr(), comm.handle.value if comm is not None else None, self.P, int(batch),
                                int(flags), 0)
        self.last_q_index = 0

    @classmethod
    def sharded(cls, Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, comm=None, **kw):
        """The problem as THIS rank of `comm` holds it: whole groups when there are enough of them, else angles
        (shard_mode_for); a plain single-GPU problem without a communicator."""
        if comm is None:
            return cls(Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, **kw)
        G = np.asarray(total).shape[1]
        if shard_mode_for(G, comm.world) == "group":
            return cls(Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, comm=comm,
                       group_range_=group_range(G, comm.rank, comm.world), **kw)
        return cls(Z, hx, mu, weight, mat, total, scat, chi, nusf, invspeed, comm=comm,
                   angles=shard_angles(np.asarray(mu, dtype=np.float64), comm.rank, comm.world), **kw)

    @classmethod
    def for_sweep(cls, Z, hx, mu, weight, sigT):
        """Bare problem for single sweeps with an explicit per-cell sigT (Z,G) (group.py:41-43)."""
        sigT = np.asarray(sigT, dtype=np.float64)
        G = sigT.shape[1]
        dev = cls(Z, hx, mu, weight, np.zeros(Z, dtype=np.int64), np.zeros((1, G)), np.zeros((1, G, G)),
                  np.zeros((1, G)), np.zeros((1, G)), np.zeros((1, G)))
        dev.t_hs = dev._upload(layout.to_storage((0.5 * sigT).T, dev.Z))
        dev.slab.hs = dev.t_hs.data_ptr()
        dev._build_tile_info()
        return dev

    # ---- plumbing ------------------------------------------------------------------------
    def _build_tile_info(self):
        """Uniform-tile records of the current hs / cdt arrays (ae_tile_info): tiles whose cells share one sigT
        take the sweep's constant-coefficient path."""
        self.t_tinfo = self.torch.empty(2 * self.G * (self.Zp // _lib.AE_TILE), dtype=self.torch.float64, device=self.dev)
        _lib.check(self.lib.ae_tile_info(self.t_hs.data_ptr(), self.t_cdt.data_ptr() if self.t_cdt is not None else None,
                                         self.Z, self.Zp, self.G, self.t_tinfo.data_ptr(), self.stream))
        self.slab.tile_info = self.t_tinfo.data_ptr()

    @property
    def phi(self):
        """[G][Zp] newest scalar-flux iterate.  The C side ping-pongs two buffers and keeps work.phi on
        the newest one; this returns the tensor that pointer refers to."""
        a, b = self._phi_bufs
        return a if self.work.phi == a.data_ptr() else b

    def _upload(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    @property
    def stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def cells_to_device(self, a, out=None):
        """(Z,G) logical ndarray / host tensor -> [G][Zp] storage tensor (packed on the device)."""
        torch = self.torch
        if not torch.is_tensor(a):
            a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(self.Z, self.G)))
        d = a.to(self.dev, non_blocking=True)
        if out is None:
            out = torch.empty(self.G * self.Zp, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_pack_cells(d.data_ptr(), self.Z, self.Zp, self.G, out.data_ptr(), self.stream))
        return out

    def cells_to_host(self, t, out=None):
        """[G][Zp] storage tensor -> (Z,G) logical ndarray (or into the pinned host tensor `out`)."""
        torch = self.torch
        d = torch.empty(self.Z * self.G, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_unpack_cells(t.data_ptr(), self.Z, self.Zp, self.G, d.data_ptr(), self.stream))
        if out is not None:
            out.copy_(d.view(out.shape), non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            return out
        return d.cpu().numpy().reshape(self.Z, self.G)

    # ---- cell-sharded host I/O (several GPUs): this rank feeds / reads only its own cell range ------------
    def cell_range(self):
        """(first logical cell, logical cells, storage offset, storage cells) of this rank's share of a row."""
        world = self.comm.world if (self.comm is not None and self.gl == 0) else 1
        rank = self.comm.rank if (self.comm is not None and self.gl == 0) else 0
        zs = self.Zp // world
        z0 = rank * zs
        return z0, max(0, min(self.Z, z0 + zs) - z0), z0, zs

    def shard_to_device(self, shard_host, rows):
        """shard_host: pinned (count, G) tensor with this rank's cells -> its range of `rows` [G][Zp], then the
        ranges of all ranks are all-gathered over NVLink so that every rank holds whole rows."""
        z0, count, s0, cells = self.cell_range()
        d = shard_host.to(self.dev, non_blocking=True)
        _lib.check(self.lib.ae_pack_cells_range(d.data_ptr(), count, cells, self.Zp, self.G, rows.data_ptr() + 8 * s0,
                                                self.stream))
        if self.comm is not None:
            _lib.check(self.lib.ae_comm_gather_rows(self.comm.handle, rows.data_ptr(), self.G, self.Zp, self.stream))

    def shard_to_host(self, rows, shard_host):
        """This rank's cell range of `rows` [G][Zp] -> pinned (count, G) tensor (synchronises)."""
        torch = self.torch
        z0, count, s0, cells = self.cell_range()
        d = torch.empty(max(count, 1) * self.G, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_unpack_cells_range(rows.data_ptr() + 8 * s0, count, cells, self.Zp, self.G, d.data_ptr(),
                                                  self.stream))
        shard_host.copy_(d[:count * self.G].view(shard_host.shape), non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()

    def psi_to_device(self, psi):
        """(Z, A_total, G) logical, reference angle order -> [Gl][A][Zp] for the local angles and groups."""
        torch = self.torch
        a = np.asarray(psi, dtype=np.float64).reshape(self.Z, self.A_total, self.G)[:, :, self.g0:self.g0 + self.Gl]
        h = torch.from_numpy(np.ascontiguousarray(a))
        d = h.to(self.dev)
        out = torch.empty(self.Gl * self.A * self.Zp, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_pack_psi(d.data_ptr(), self.Z, self.Zp, self.A_total, self.Gl, self.t_aidx.data_ptr(), self.A,
                                        out.data_ptr(), self.stream))
        return out

    def psi_to_host(self, t):
        """[Gl][A][Zp] -> (Z, A, Gl) for the local angles (device angle order) and groups."""
        torch = self.torch
        d = torch.zeros(self.Z * self.A_total * self.Gl, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_unpack_psi(t.data_ptr(), self.Z, self.Zp, self.A_total, self.Gl, self.t_aidx.data_ptr(), self.A,
                                          d.data_ptr(), self.stream))
        full = d.cpu().numpy().reshape(self.Z, self.A_total, self.Gl)
        return np.ascontiguousarray(full[:, self.angles, :])

    def new_psi(self):
        return self.torch.zeros(self.Gl * self.A * self.Zp, dtype=self.torch.float64, device=self.dev)

    # ---- C entry points --------------------------------------------------------------------
    def sweep_set(self, q, psi_in=None, psi_out=None):
        """ae_sweep_set + ae_reduce_partials: returns the [G][Zp] scalar flux of the local angles
        (summed over ranks when a communicator is attached)."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.lib.ae_sweep_set(C.byref(self.slab), q.data_ptr(), ptr(psi_in), ptr(psi_out),
                                         self.partial.data_ptr(), self.sync.data_ptr(), self.stream))
        phi = self.torch.empty(self.G * self.Zp, dtype=self.torch.float64, device=self.dev)
        _lib.check(self.lib.ae_reduce_partials(C.byref(self.slab), self.partial.data_ptr(), self.P, phi.data_ptr(),
                                               self.stream))
        self._complete_flux(phi)
        return phi

    def _complete_flux(self, phi):
        """Make a locally reduced flux [G][Zp] whole: sum over ranks (angle-sharded) or gather the ranks' groups."""
        if self.comm is None:
            return
        if self.gl > 0:
            _lib.check(self.lib.ae_comm_gather_groups(self.comm.handle, phi.data_ptr(), self.G, self.Zp, self.stream))
        else:
            self.comm.allreduce_(phi)

    def sweep_set_phi(self, q):
        """ae_sweep_set_phi + ae_reduce_partials: scalar flux [G][Zp] of a sweep set without angular flux in or out
        (the kernel the solver loops use for their phi-only iterations; it may write fewer partial slabs)."""
        P = C.c_int32(0)
        _lib.check(self.lib.ae_sweep_set_phi(C.byref(self.slab), q.data_ptr(), self.partial.data_ptr(), self.sync.data_ptr(),
                                             C.byref(P), self.stream))
        phi = self.torch.empty(self.G * self.Zp, dtype=self.torch.float64, device=self.dev)
        _lib.check(self.lib.ae_reduce_partials(C.byref(self.slab), self.partial.data_ptr(), P.value, phi.data_ptr(),
                                               self.stream))
        self._complete_flux(phi)
        self.last_phi_partials = P.value
        return phi

    def flux_update(self, partial, P, q_next, tol=0.0, iteration=0):
        """ae_flux_update: phi assembly + norm + next scatter source (w->base must be set)."""
        _lib.check(self.lib.ae_flux_update(C.byref(self.slab), C.byref(self.work), partial.data_ptr(), P,
                                           q_next.data_ptr(), tol, iteration, self.stream))

    def sweep_and_update(self, q, psi_in, psi_out, q_next, tol=0.0, iteration=0):
        """ae_sweep_and_update: one sweep set + flux update as one C call (with AE_COMM_PEER_COPY=1 and AE_EXCHANGE_OVERLAP=1
        group-sharded ranks split it in two group chunks and exchange the first behind the sweep of the second)."""
        ptr = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.lib.ae_sweep_and_update(C.byref(self.slab), C.byref(self.work), q.data_ptr(), ptr(psi_in), ptr(psi_out),
                                                q_next.data_ptr(), tol, iteration, self.stream))

    def read_ctrl(self):
        """Copy the device control block to the host (synchronises the stream)."""
        self.ctrl_host.copy_(self.ctrl, non_blocking=True)
        self.torch.cuda.current_stream(self.dev).synchronize()
        return _lib.AeCtrl.from_buffer_copy(bytes(self.ctrl_host.numpy()))

    def source_iteration(self, base, psi_in=None, cold_start=1e-12, tol=1e-8, max_iterations=100):
        """ae_source_iteration.  `base` [G][Zp] tensor is the fixed source.  Returns (status, iterations)."""
        self.base.copy_(base)
        it, lq = C.c_int32(0), C.c_int32(0)
        st = self.lib.ae_source_iteration(C.byref(self.slab), C.byref(self.work),
                                          psi_in.data_ptr() if psi_in is not None else None, cold_start, tol,
                                          max_iterations, C.byref(it), C.byref(lq), self.stream)
        self.last_q_index = lq.value
        return st, it.value

    def power_iteration(self, phi_old, tol=1e-8, max_iterations=100, si_tol=1e-8, si_max_iterations=100,
                        psi_out=None):
        """ae_power_iteration.  phi_old: (Z,G) ndarray (phi.old).  Returns dict like the oracle's."""
        self.phi_outer.copy_(self.cells_to_device(phi_old))
        k_new, k_old, it = C.c_double(0), C.c_double(0), C.c_int32(0)
        k_hist = (C.c_double * max_iterations)()
        si_hist = (C.c_int32 * max_iterations)()
        lq = C.c_int32(0)
        st = self.lib.ae_power_iteration(C.byref(self.slab), C.byref(self.work), tol, max_iterations, si_tol,
                                         si_max_iterations, psi_out.data_ptr() if psi_out is not None else None,
                                         C.byref(k_new), C.byref(k_old), C.byref(it), k_hist, si_hist, C.byref(lq),
                                         self.stream)
        self.last_q_index = lq.value
        n = it.value
        return dict(status=st, k_new=k_new.value, k_old=k_old.value, iterations=n,
                    phi_new=self.cells_to_host(self.phi), phi_old=self.cells_to_host(self.phi_outer),
                    k_hist=np.array(k_hist[:n]), si_hist=np.array(si_hist[:n], dtype=np.int32))

    def time_step(self, psi_prev, psi_next, tol=1e-10):
        outer, inner = C.c_int32(0), C.c_int32(0)
        st = self.lib.ae_time_step(C.byref(self.slab), C.byref(self.work), psi_prev.data_ptr(), psi_next.data_ptr(),
                                   tol, C.byref(outer), C.byref(inner), self.stream)
        return st, outer.value, inner.value

    def time_march(self, psi, num_steps, tol=1e-10, psi_snap=None, phi_snap=None):
        """ae_time_march.  psi: [G][A][Zp] tensor (psi0 in, last psi out)."""
        outer = (C.c_int32 * num_steps)()
        inner = (C.c_int32 * num_steps)()
        st = self.lib.ae_time_march(C.byref(self.slab), C.byref(self.work), psi.data_ptr(),
                                    psi_snap.data_ptr() if psi_snap is not None else None,
                                    phi_snap.data_ptr() if phi_snap is not None else None, num_steps, tol,
                                    outer, inner, self.stream)
        return st, np.array(outer[:], dtype=np.int32), np.array(inner[:], dtype=np.int32)

    # ---- DMD ---------------------------------------------------------------------------------
    def gram(self, snap, row_len, first, count, sharded=True):
        """Gram matrix [count][count] of snapshot rows first..first+count-1 of `snap` [steps][row_len];
        summed over ranks when the rows are sharded."""
        g = self.torch.empty(count * count, dtype=self.torch.float64, device=self.dev)
        _lib.check(self.lib.ae_gram(snap.data_ptr() + 8 * first * row_len, row_len, row_len, count, g.data_ptr(),
                                    self.stream))
        if self.comm is not None and sharded:
            self.comm.allreduce_(g)
        return g

    def tsqr(self, snap, row_len, first, count, sharded=True):
        """Triangular factor R [count][count] of the snapshot window (Householder TSQR, Q never formed).
        Row-sharded ranks (`sharded`: every rank holds different rows, e.g. the psi rows of its angles)
        exchange their factors (allreduce of a zero-padded stack) and merge them."""
        torch = self.torch
        r = torch.empty(count * count, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.ae_tsqr(snap.data_ptr() + 8 * first * row_len, row_len, row_len, count, r.data_ptr(),
                                    self.stream))
        if self.comm is not None and sharded:
            stack = torch.zeros(self.comm.world * count * count, dtype=torch.float64, device=self.dev)
            stack[self.comm.rank * count * count:(self.comm.rank + 1) * count * count] = r
            self.comm.allreduce_(stack)
            _lib.check(self.lib.ae_tsqr_merge(stack.data_ptr(), self.comm.world, count, r.data_ptr(), self.stream))
        return r

    def dmd_from_r(self, r, K, dt, rank_tol=1e-13):
        are, aim, sig = (C.c_double * K)(), (C.c_double * K)(), (C.c_double * K)()
        rk = C.c_int32(0)
        _lib.check(self.lib.ae_dmd_from_r(r.data_ptr(), K, dt, rank_tol, are, aim, sig, C.byref(rk), self.stream))
        n = rk.value
        return np.array(are[:n]) + 1j * np.array(aim[:n]), np.array(sig[:]), n

    def dmd_from_r_batched(self, r_stack, count, K, dt, rank_tol=1e-13):
        """ae_dmd_from_r_batched: `count` triangular factors [count][K+1][K+1] (device) -> list of (alphas, sigma, rank)
        or the status code of a window that failed."""
        are, aim, sig = (np.zeros((count, K)) for _ in range(3))
        rk, st = np.zeros(count, dtype=np.int32), np.zeros(count, dtype=np.int32)
        _lib.check(self.lib.ae_dmd_from_r_batched(r_stack.data_ptr(), count, K, dt, rank_tol, are.ctypes.data, aim.ctypes.data,
                                                  sig.ctypes.data, rk.ctypes.data, st.ctypes.data, self.stream))
        return [(are[m, :rk[m]] + 1j * aim[m, :rk[m]], sig[m].copy(), int(rk[m])) if st[m] == 0 else int(st[m])
                for m in range(count)]

    def dmd_from_gram(self, gram, K, dt, rank_tol=1e-13):
        are, aim, sig = (C.c_double * K)(), (C.c_double * K)(), (C.c_double * K)()
        r = C.c_int32(0)
        _lib.check(self.lib.ae_dmd_from_gram(gram.data_ptr(), K, dt, rank_tol, are, aim, sig, C.byref(r), self.stream))
        n = r.value
        alphas = np.array(are[:n]) + 1j * np.array(aim[:n])
        return alphas, np.array(sig[:]), n
