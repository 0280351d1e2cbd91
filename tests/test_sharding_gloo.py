"""world_size-2 CPU tests of the two decompositions of a sweep set: angle shards summed with
a gloo allreduce, group shards gathered with gloo broadcasts, both equal to the single-process flux; plus the host
logic that picks the decomposition and the group ranges (against the library's own ae_group_range)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from alpha_eigen_b200.device import group_range, shard_angles, shard_mode_for


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = kp_problem(9, 10, 8)
    src = np.random.RandomState(5).rand(p["Z"], 1) + 0.1
    mine = shard_angles(p["mu"], rank, world)
    phi = orc.sweep_set(p["mu"][mine], p["weight"][mine], p["hx"], src, p["total_xs"])
    t = torch.from_numpy(phi.copy())
    dist.all_reduce(t)
    full = orc.sweep_set(p["mu"], p["weight"], p["hx"], src, p["total_xs"])
    q.put((rank, float(np.max(np.abs(t.numpy() - full) / np.abs(full))), sorted(mine.tolist())))
    dist.destroy_process_group()


def test_two_rank_angle_shards_sum_to_full_flux():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    owned = sorted(sum((o[2] for o in out), []))
    assert owned == list(range(8))                      # every angle owned exactly once
    for _, err, mine in out:
        assert err < 1e-14
        assert len(mine) == 4


def test_shard_angles_balanced():
    mu, _ = np.polynomial.legendre.leggauss(256)
    seen = []
    for r in range(8):
        a = shard_angles(mu, r, 8)
        assert len(a) == 32 and np.sum(mu[a] < 0) == 16
        assert np.all(mu[a][:16] < 0) and np.all(mu[a][16:] > 0)
        seen += a.tolist()
    assert sorted(seen) == list(range(256))


def _group_worker(rank, world, port, q):
    import torch
    from alpha_eigen_b200.synthetic import synthetic_problem
    from oracle import ae_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synthetic_problem(300, 8, 7)
    src = np.random.RandomState(5).rand(p["Z"], p["G"]) + 0.1
    sigT = p["total"][p["mat"]]
    g0, gl = group_range(p["G"], rank, world)
    rows = torch.zeros(p["G"], p["Z"], dtype=torch.float64)
    # this rank sweeps ITS groups with ALL angles: their flux is complete, nothing to reduce
    mine = orc.sweep_set(p["mu"], p["weight"], p["hx"], src[:, g0:g0 + gl], sigT[:, g0:g0 + gl])
    rows[g0:g0 + gl] = torch.from_numpy(np.ascontiguousarray(mine.T))
    for r in range(world):                              # the all-gather of unequal blocks: one broadcast per owner
        a, n = group_range(p["G"], r, world)
        blk = rows[a:a + n].contiguous()
        dist.broadcast(blk, r)
        rows[a:a + n] = blk
    full = orc.sweep_set(p["mu"], p["weight"], p["hx"], src, sigT)
    q.put((rank, float(np.max(np.abs(rows.numpy().T - full) / np.abs(full))), (g0, gl)))
    dist.destroy_process_group()


def test_two_rank_group_shards_gather_to_full_flux():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_group_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [o[2] for o in out] == [(0, 4), (4, 3)]       # 7 groups over 2 ranks
    for _, err, _ in out:
        assert err == 0.0                                # same sweeps, only placed by another rank


def test_group_ranges_match_library_and_cover_every_group():
    import ctypes as C
    from alpha_eigen_b200 import _lib
    lib = _lib.load()
    for G, world in [(70, 8), (70, 4), (70, 2), (12, 5), (16, 8), (9, 3), (7, 2)]:
        seen = []
        for r in range(world):
            g0, gl = group_range(G, r, world)
            a, b = C.c_int32(-1), C.c_int32(-1)
            lib.ae_group_range(G, world, r, C.byref(a), C.byref(b))
            assert (a.value, b.value) == (g0, gl)
            seen += list(range(g0, g0 + gl))
        assert seen == list(range(G))
        counts = [group_range(G, r, world)[1] for r in range(world)]
        assert max(counts) - min(counts) <= 1


def test_decomposition_choice():
    assert shard_mode_for(70, 1) == "single"
    assert [shard_mode_for(70, n) for n in (2, 4, 8)] == ["group"] * 3     # BASELINE configs[3]
    assert shard_mode_for(1, 2) == "angle" and shard_mode_for(1, 8) == "angle"   # Kornreich-Parsons
    assert shard_mode_for(12, 8) == "angle" and shard_mode_for(12, 4) == "group"   # ensemble members, row-sharded
    assert shard_mode_for(8, 2) == "angle"                                # one-thread-per-cell update kernel: cells are sharded
