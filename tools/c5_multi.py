#!/usr/bin/env python
"""BASELINE configs[4] on the whole box: the 1024-member ensemble of 12-group slabs with DMD, under torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        tools/c5_multi.py [--members 1024] [--row-sharded 8]

Two lines of JSON (rank 0): member-parallel (members dealt over the ranks, batched march kernel + DMDs, no communication
until the results are gathered) and row-sharded (every member solved by ALL ranks: angle x group rows dealt over the
ranks, psi snapshot matrix sharded by rows, per-rank TSQR factors merged -- the layout BASELINE.json names), the latter on
a sample of members, one of which is checked against its single-GPU spectrum.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1024)
    ap.add_argument("--row-sharded", type=int, default=8, help="members solved in row-sharded mode")
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from alpha_eigen_b200 import distributed
    from alpha_eigen_b200.ensemble import Ensemble, solve_member, synthetic_ensemble
    comm = distributed.init()
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    Z, A, G = 2048, 16, 12
    members = synthetic_ensemble(1024, Z, A, G)[:: max(1024 // args.members, 1)][:args.members]
    fence = (lambda: (dist.barrier(), torch.cuda.synchronize())) if world > 1 else torch.cuda.synchronize
    srt = lambda a: np.asarray(a)[np.lexsort((np.asarray(a).imag, -np.asarray(a).real))]

    Ensemble(members[:2 * world], dt=0.1, num_steps=10).run()                       # warm-up (contexts, modules)
    fence(); t0 = time.perf_counter()
    res = Ensemble(members, dt=0.1, num_steps=args.steps).run()
    fence(); t1 = time.perf_counter()
    sweeps = sum(int(np.sum(r["inner"])) + args.steps for r in res)
    if rank == 0:
        print(json.dumps({"case": "c5_member_parallel", "gpus": world, "members": len(members), "Z": Z, "A": A, "G": G,
                          "steps": args.steps, "wall_s": t1 - t0, "members_per_s": len(members) / (t1 - t0),
                          "sweep_sets": sweeps, "updates_per_s": float(sweeps) * Z * A * G / (t1 - t0),
                          "all_converged": all(r["status"] == 0 for r in res),
                          "alpha0_first_last": [float(srt(res[0]["alphas"])[0].real), float(srt(res[-1]["alphas"])[0].real)],
                          "ranks": [int(res[0]["rank"]), int(res[-1]["rank"])]}), flush=True)
    if world > 1 and args.row_sharded > 0:
        sample = members[:args.row_sharded]
        solve_member(sample[0], 0.1, 10, comm=comm)                                # warm-up
        fence(); t0 = time.perf_counter()
        rs = [solve_member(p, 0.1, args.steps, comm=comm) for p in sample]
        fence(); t1 = time.perf_counter()
        a_row, a_one = srt(rs[0]["alphas"]), srt(res[0]["alphas"])
        same_iters = bool(np.array_equal(rs[0]["inner"], res[0]["inner"]) and np.array_equal(rs[0]["outer"], res[0]["outer"]))
        if rank == 0:
            print(json.dumps({"case": "c5_row_sharded", "gpus": world, "members": len(sample), "steps": args.steps,
                              "wall_s": t1 - t0, "members_per_s": len(sample) / (t1 - t0),
                              "sharding": "groups" if G >= 2 * world and G > 8 else "angles",
                              "member0_vs_single_gpu": {"iterations_equal": same_iters, "rank_row": int(rs[0]["rank"]),
                                                        "rank_single": int(res[0]["rank"]),
                                                        "alpha0_rel_diff": float(abs(a_row[0] - a_one[0]) / abs(a_one[0])),
                                                        "sigma_rel_diff": float(np.max(np.abs(rs[0]["sigma"] - res[0]["sigma"])) / res[0]["sigma"][0])}}),
                  flush=True)
    distributed.shutdown()


if __name__ == "__main__":
    main()
