#!/usr/bin/env python
"""Benchmark of the slab S_N hot path: sweep cell*angle*group updates/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c3|kp]

One "step" = one time-dependent source iteration over the whole slab: a psi-streaming sweep set
(read psi^n, write psi^{n+1} in place, accumulate the scalar flux; group.py:41-60,74-77 with the
angular source of time_unit.py:23), for N > 1 the reduce-scatter of the partial flux rows over NVLink,
the flux update kernels (norm + scatter source; group.py:73,77-80) on this rank's 1/N of the cells and the
all-gather of the new source.  Rows (angle x group) are sharded by angle over the ranks; every N runs the
SAME named problem (strong scaling).

Default workload c4 = BASELINE.json configs[3]: synthetic 70-group slab, S256, 1M cells
(17 920 rows x 1e6 cells = 1.79e10 updates per step, psi = 143 GB updated in place).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Z, A, G, description)
    "c4": (1_000_000, 256, 70, "synthetic 70-group slab, S256, 1M cells (BASELINE configs[3])"),
    "c3": (100_000, 64, 70, "synthetic 70-group slab, S64, 100k cells (BASELINE configs[2])"),
    "kp": (1800, 196, 1, "Kornreich-Parsons long: 200 cells/mfp, S196, 1 group (BASELINE configs[1])"),
}
DT = 0.1                                # inputs.py:39


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(seconds=12.0):
    """Oracle port (reference arithmetic, group.py:41-60 + time_unit.py:23) timed on the host cores
    on a bounded slice of the same kind of work: time-dependent sweep sets of a 70-group slab."""
    from alpha_eigen_b200.synthetic import synthetic_problem
    from oracle import ae_oracle as orc
    Z, A, G = 20_000, 16, 70
    p = synthetic_problem(Z, A, G)
    rs = np.random.RandomState(1)
    inv = p["invspeed"][p["mat"]]
    sigT = p["total"][p["mat"]] + 1 / DT * inv
    q = rs.rand(Z, G)
    psi = rs.rand(Z, A, G)
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    orc.bench_sweep_sets(p["mu"], p["weight"], p["hx"], q, sigT, inv / DT, psi, 1)       # warm-up
    reps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        orc.bench_sweep_sets(p["mu"], p["weight"], p["hx"], q, sigT, inv / DT, psi, 1)
        reps += 1
    el = time.perf_counter() - t0
    return {"value": Z * A * G * reps / el, "unit": "cell*angle*group updates/s", "cores": min(cores, G),
            "kind": "port",
            "sample": f"{reps} time-dependent sweep sets of a synthetic {G}-group slab, Z={Z}, S{A} "
                      f"({Z * A * G * reps:.3g} updates, {el:.1f} s); literal Python reference measured at "
                      f"1.6-1.75e5 updates/s on 1 core (BASELINE.md)"}


def run_reference(args, Z, A, G, desc):
    """--impl reference: the reference algorithm on the host cores (oracle port: the Python reference
    cannot be instantiated at these sizes, one_direction.py:39 allocates (Z+1)^2 doubles per angle)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 12.0 / max(args.steps + args.warmup, 1)
    t0 = time.perf_counter()
    cb = cpu_baseline(seconds=max(per_step * (args.steps + args.warmup), 6.0))
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps + args.warmup, 1)
    line = {"impl": "reference", "metric": "sweep cell*angle*group updates/s", "value": cb["value"],
            "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "Z": Z, "A": A, "G": G}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=0, help="override Z (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    Z, A, G, desc = WORKLOADS[args.workload]
    if args.cells:
        Z = args.cells
        desc += f" [Z overridden to {Z}]"
    if args.impl == "reference":
        return run_reference(args, Z, A, G, desc)

    import torch
    import torch.distributed as dist
    from alpha_eigen_b200 import _lib, distributed
    from alpha_eigen_b200.device import HostPipe, SlabDevice, shard_angles
    from alpha_eigen_b200.synthetic import synthetic_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = distributed.init() if world > 1 else None
    W = max(args.warmup, 3)

    p = synthetic_problem(Z, A, G)
    angles = shard_angles(p["mu"], rank, world)
    dev = SlabDevice(Z, p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                     p["invspeed"], dt=DT, angles=angles, comm=comm)
    n_phi = dev.G * dev.Zp
    psi = torch.empty(dev.G * dev.A * dev.Zp, dtype=torch.float64, device=dev.dev)
    psi.uniform_(0.5, 1.5, generator=torch.Generator(device=dev.dev).manual_seed(2021 + rank))
    # experiment knob: out-of-place psi update (needs a second psi buffer; the headline run is in place)
    psis = [psi, torch.empty_like(psi)] if os.environ.get("AE_BENCH_PSI_PINGPONG") == "1" else None
    dev.base.zero_()
    q = [dev.q0, dev.q1]
    q[0].uniform_(0.0, 1.0, generator=torch.Generator(device=dev.dev).manual_seed(7))
    updates_per_step = Z * A * G                        # whole job, all ranks
    local_updates = Z * dev.A * G
    launches = 0

    def step(i, ev=None):
        nonlocal launches
        if ev is not None:
            ev[0].record()
        pin, pout = (psis[i & 1], psis[(i + 1) & 1]) if psis is not None else (psi, psi)
        dev.lib.ae_sweep_set(dev.slab, q[i & 1].data_ptr(), pin.data_ptr(), pout.data_ptr(), dev.partial.data_ptr(),
                             dev.sync.data_ptr(), dev.stream)
        if ev is not None:
            ev[1].record()
        launches += 1
        if comm is not None:
            _lib.check(dev.lib.ae_exchange_partials(dev.slab, dev.work, dev.stream))   # local sum + reduce-scatter by rows
            if ev is not None:
                ev[2].record()
            dev.flux_update(dev.reduced, 1, q[(i + 1) & 1], 0.0, i + 1)             # 1/N of the cells + all-gather of q
            launches += 5                                   # reduce, (nccl), assemble, scatter, (nccl), norm
        else:
            if ev is not None:
                ev[2].record()
            dev.flux_update(dev.partial, dev.P, q[(i + 1) & 1], 0.0, i + 1)
            launches += 2                                   # assemble, scatter
        if ev is not None:
            ev[3].record()

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    fence()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = 0
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    t_beg.record()
    for i in range(args.steps):
        step(W + i, evs[i])
    t_end.record()
    fence()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_beg.elapsed_time(t_end)
    phase = [float(np.mean([e[k].elapsed_time(e[k + 1]) for e in evs])) for k in range(3)]
    if world > 1:
        t = torch.tensor([ms_total] + phase, dtype=torch.float64, device=dev.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, phase = float(t[0]), [float(x) for x in t[1:]]
    sweep_ms = phase[0]
    ms_per_step = ms_total / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)
    gpu_launches = launches
    assert bool(torch.isfinite(dev.phi).all()), "non-finite scalar flux"

    # ---- end to end through host buffers: the source comes from pinned host memory, the flux goes back ---------
    # Several GPUs: host I/O is sharded by cells like the flux update -- every rank uploads the source of ITS cell
    # range (1/N of the array), the ranges are all-gathered over NVLink, and every rank reads back its range of the flux.
    z0, zcount, _, _ = dev.cell_range()
    q_host = torch.rand(max(zcount, 1), G, dtype=torch.float64)[:zcount].pin_memory()
    phi_host = torch.empty(zcount, G, dtype=torch.float64).pin_memory()
    e_steps = max(3, min(args.steps, 8))

    # Every step still uploads its own source and downloads its own flux; the copies run on the copy engines
    # (HostPipe: two staging slots, own streams) so that the upload of step i+1 and the download of step i-1
    # overlap the kernels of step i.  AE_BENCH_E2E_SERIAL=1 times the unpipelined sequence instead.
    serial = os.environ.get("AE_BENCH_E2E_SERIAL") == "1"
    pipe = HostPipe(dev)

    def e2e_run(n):
        if serial:
            for i in range(n):
                dev.shard_to_device(q_host, q[i & 1])                   # H2D + pack (+ all-gather of the ranges)
                step(i)
                dev.shard_to_host(dev.phi, phi_host)                    # unpack + D2H (+ sync)
            return
        pipe.upload(0, q_host)
        for i in range(n):
            pipe.feed(i, q[i & 1])                                      # pack (+ all-gather of the ranges)
            if i + 1 < n:
                pipe.upload(i + 1, q_host)                              # H2D of the next source, behind the sweep
            step(i)
            pipe.drain(i, dev.phi, phi_host)                            # unpack, D2H behind the next sweep
        pipe.join()                                                     # the last flux has reached the host

    e2e_run(2)
    fence()
    pipe.wait()
    t_beg.record()
    e2e_run(e_steps)
    t_end.record()
    fence()
    pipe.wait()
    e_ms = t_beg.elapsed_time(t_end) / e_steps
    if world > 1:
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t[0])

    if rank == 0:
        hbm_peak, peak_src = peaks()
        bytes_per_update = 16.0 + 24.0 / dev.A + 8.0 * dev.P / dev.A    # psi r+w, hs/q/cdt over angles, partial
        alg_bytes = bytes_per_update * local_updates
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one sweep launch (profiles/r01_c4_sweep_dram_bytes_v6_gangs.csv:
        # 152.7 + 152.3 GB; 170.1 + 152.3 GB before the gang order, profiles/r01_c4_kernel_dram_bytes_v5.csv)
        traffic = 305.0e9 if (args.workload == "c4" and world == 1 and not args.cells) else None
        line = {
            "metric": "sweep cell*angle*group updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "Z": Z, "A": A, "G": G, "dt": DT, "rows_per_gpu": dev.A * G,
                       "sharding": f"angle x group rows dealt by angle over {world} rank(s)",
                       "step": "psi-streaming sweep set (in place) + flux reduce-scatter (N>1) + flux-update kernels on "
                               "1/N of the cells + all-gather of the new source (N>1)",
                       "l2": "inputs (psi %.1f GB per GPU) are larger than L2" % (psi.numel() * 8 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "algorithmic_bytes": alg_bytes, "kernel": "ae::sweep_kernel<16,true,true> (8 for fewer than 16 angles per direction)",
                         "kernel_ms": sweep_ms, "bytes_per_update": bytes_per_update, "peak_source": peak_src,
                         "sweep_share_of_step": sweep_ms / ms_per_step},
            "breakdown_ms": {"sweep": phase[0], "exchange_reduce_scatter": phase[1],
                             "flux_update_and_allgather": phase[2]},
            "e2e": {"value": updates_per_step / (e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e_ms,
                    "h2d_bytes_per_step": Z * G * 8, "d2h_bytes_per_step": Z * G * 8,
                    "steps": e_steps, "pipelined": not serial,
                    "what": "whole-job bytes; every step every rank copies the source of its own 1/N cell range from pinned "
                            "host memory (all-gathered on the device), runs the step on resident psi, and reads its range of "
                            "the flux back to pinned host memory; copies run on their own streams, double-buffered, so the "
                            "upload of step i+1 and the download of step i-1 overlap the kernels of step i"},
            "gpu_launches": gpu_launches, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        distributed.shutdown()


if __name__ == "__main__":
    main()
