"""Multi-GPU parity check, run under torchrun on N GPUs of one box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Angle-sharded (one-group K-P) and group-sharded (16-group synthetic slab) power iteration, time march and DMD
(row-sharded TSQR + Gram) against the single-process oracle: identical iteration counts, k to 1e-9, flux to 1e-10,
alpha to the single-GPU answer.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from alpha_eigen_b200 import distributed
    from alpha_eigen_b200.device import SlabDevice, shard_angles
    from alpha_eigen_b200.time_unit import dmd_window
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem, kp_psi0
    comm = distributed.init()
    rank, world = comm.rank, comm.world
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

    p = kp_problem(9, 20, 8)
    t = p["tables"]
    mine = shard_angles(p["mu"], rank, world)
    dev = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], t.mat, t.total, t.scat, t.chi, t.nusf, t.invspeed,
                     angles=mine, comm=comm)
    rd = dev.power_iteration(p["phi0"])
    ro = orc.power_iteration(p["mu"], p["weight"], p["hx"], t, p["phi0"])
    assert rd["status"] == 0 and list(rd["si_hist"]) == list(ro["si_hist"]), (rd["si_hist"], ro["si_hist"])
    assert abs(rd["k_new"] - ro["k_new"]) < 1e-9
    assert rel(rd["phi_new"], ro["phi_new"]) < 1e-10

    steps = 20
    devt = SlabDevice(p["Z"], p["hx"], p["mu"], p["weight"], t.mat, t.total, t.scat, t.chi, t.nusf, t.invspeed,
                      dt=0.1, angles=mine, comm=comm)
    psi = devt.psi_to_device(kp_psi0(p))
    n_psi = devt.G * devt.A * devt.Zp
    snap = torch.zeros(steps * n_psi, dtype=torch.float64, device=devt.dev)
    st, outer, inner = devt.time_march(psi, steps, 1e-10, snap, None)
    res = orc.time_march(p["mu"], p["weight"], p["hx"], t, np.zeros((p["Z"], 1)), kp_psi0(p), 0.1, steps)
    assert st == 0 and list(outer) == list(res["outer"]) and list(inner) == list(res["inner"])
    assert rel(devt.psi_to_host(psi), res["psi_last"][:, mine, :]) < 1e-10
    # row-sharded DMD: every rank holds the rows of its own angles
    first, K = dmd_window(steps)
    R = devt.tsqr(snap, n_psi, first, K + 1)
    al, sig, r = devt.dmd_from_r(R, K, 0.1)
    ref = orc.dmd_alpha_from_psi(res["psi_snap"], steps, 0.1)
    srt = lambda a: np.asarray(a)[np.lexsort((np.asarray(a).imag, -np.asarray(a).real))]
    S = res["psi_snap"].reshape(-1, steps)[:, first:first + K]
    sv = np.linalg.svd(S, compute_uv=False)
    assert np.max(np.abs(sig[:len(sv)] - sv)) < 1e-14 * sv[0], np.max(np.abs(sig[:len(sv)] - sv)) / sv[0]
    assert r == len(ref), (r, len(ref))
    dom_err = abs(srt(al)[0] - srt(ref)[0]) / abs(srt(ref)[0])
    G = devt.gram(snap, n_psi, first, K + 1).cpu().numpy().reshape(K + 1, K + 1)
    Sf = res["psi_snap"].reshape(-1, steps)[:, first:first + K + 1]
    assert np.max(np.abs(G - Sf.T @ Sf)) < 1e-13 * np.max(np.abs(Sf.T @ Sf))
    # ---- group-sharded: 16 groups, S64 (the rows-per-warp phi kernel), streaming kernels ---------------------------
    from alpha_eigen_b200.synthetic import synthetic_problem
    q = synthetic_problem(1500, 64, 16)
    qt = orc.Tables(q["mat"], q["total"], q["scat"], q["chi"], q["nusf"], q["invspeed"])
    rs = np.random.RandomState(5)
    half = rs.rand(q["Z"], q["G"])
    phi0 = half + half[::-1]
    mk = lambda **kw: SlabDevice.sharded(q["Z"], q["hx"], q["mu"], q["weight"], qt.mat, qt.total, qt.scat, qt.chi, qt.nusf,
                                         qt.invspeed, comm=comm, flags=1, **kw)
    dg = mk()
    assert dg.shard_mode == ("group" if world <= 8 else "angle"), dg.shard_mode
    rg = dg.power_iteration(phi0)
    og = orc.power_iteration(q["mu"], q["weight"], q["hx"], qt, phi0)
    assert rg["status"] == 0 and list(rg["si_hist"]) == list(og["si_hist"]), (rg["si_hist"], og["si_hist"])
    assert abs(rg["k_new"] - og["k_new"]) < 1e-9, (rg["k_new"], og["k_new"])
    assert rel(rg["phi_new"], og["phi_new"]) < 1e-10
    psi0 = rs.rand(q["Z"], q["A"], q["G"])
    s_ext = rs.rand(q["Z"], q["G"]) * 0.1
    dgt = mk(dt=0.1, s_ext=s_ext)
    pg = dgt.psi_to_device(psi0)
    gsteps = 10
    n_pg = dgt.Gl * dgt.A * dgt.Zp
    gsnap = torch.zeros(gsteps * n_pg, dtype=torch.float64, device=dgt.dev)
    st, outer, inner = dgt.time_march(pg, gsteps, 1e-10, gsnap, None)
    rm = orc.time_march(q["mu"], q["weight"], q["hx"], qt, s_ext, psi0, 0.1, gsteps)
    assert st == 0 and list(outer) == list(rm["outer"]) and list(inner) == list(rm["inner"])
    assert rel(dgt.cells_to_host(dgt.phi), rm["phi_last"]) < 1e-10
    assert rel(dgt.psi_to_host(pg), rm["psi_last"][:, dgt.angles, :][..., dgt.groups]) < 1e-10
    # ae_sweep_and_update against the two plain calls; with the peer-copy exchange (AE_COMM_PEER_COPY=1) it runs as two
    # group chunks, the exchange of the first behind the sweep of the second
    if dgt.gl > 0:
        qa = torch.rand(dgt.G * dgt.Zp, dtype=torch.float64, device=dgt.dev, generator=torch.Generator(device=dgt.dev).manual_seed(3))
        dgt.base.zero_()
        outs = []
        import os
        os.environ["AE_EXCHANGE_OVERLAP"] = "1"                         # the two-chunk step is opt-in
        for fused in (False, True):
            pa = dgt.psi_to_device(psi0)
            qn = torch.zeros_like(qa)
            for b in dgt._phi_bufs:                                     # same previous iterate for both runs
                b.fill_(0.25)
            if fused:
                dgt.sweep_and_update(qa, pa, pa, qn, 0.0, 1)
            else:
                import ctypes as C
                from alpha_eigen_b200 import _lib
                _lib.check(dgt.lib.ae_sweep_set(C.byref(dgt.slab), qa.data_ptr(), pa.data_ptr(), pa.data_ptr(),
                                                dgt.partial.data_ptr(), dgt.sync.data_ptr(), dgt.stream))
                dgt.flux_update(dgt.partial, dgt.P, qn, 0.0, 1)
            torch.cuda.synchronize()
            outs.append((dgt.phi.clone(), qn.clone(), pa.clone(), dgt.read_ctrl().change))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
        assert abs(outs[0][3] - outs[1][3]) < 1e-13 * abs(outs[0][3])
        xmode = dgt.lib.ae_comm_exchange_mode(dgt.comm.handle)
        want = 1 if os.environ.get("AE_COMM_PEER_COPY") == "1" else 0
        assert xmode == want, f"flux exchange mode {xmode}, expected {want} (1 = peer copies, 0 = NCCL)"
        if rank == 0:
            print(f"group-sharded flux exchange: {'peer copies (copy engines + stream memory operations)' if xmode == 1 else 'NCCL all-gather'}")
    gfirst, gK = dmd_window(gsteps)
    Rg = dgt.tsqr(gsnap, n_pg, gfirst, gK + 1)                       # rows of this rank's groups; factors merged across ranks
    alg, sigg, rkg = dgt.dmd_from_r(Rg, gK, 0.1)
    Sg = rm["psi_snap"].reshape(-1, gsteps)[:, gfirst:gfirst + gK]
    svg = np.linalg.svd(Sg, compute_uv=False)
    assert np.max(np.abs(sigg[:len(svg)] - svg)) < 1e-13 * svg[0], np.max(np.abs(sigg[:len(svg)] - svg)) / svg[0]
    if rank == 0:
        print(f"mgpu_check ok on {world} GPUs: k={rd['k_new']:.15f} (oracle {ro['k_new']:.15f}), "
              f"alpha0={srt(al)[0]:.12g} (reference routine {srt(ref)[0]:.12g}, rel diff {dom_err:.2e}), rank {r}; "
              f"group-sharded 16-group slab: k={rg['k_new']:.15f} (oracle {og['k_new']:.15f}), {gsteps} steps, "
              f"iterations {int(np.sum(inner))} == oracle")
    distributed.shutdown()


if __name__ == "__main__":
    main()
