"""world_size-2 CPU test of the angle sharding: per-rank partial scalar fluxes (oracle sweeps over
each rank's angles) summed with a gloo allreduce equal the single-process flux."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from alpha_eigen_b200.device import shard_angles


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
