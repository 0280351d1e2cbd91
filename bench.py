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

Besides the headline step the line carries: `roofline` (the psi-streaming sweep kernel against the measured HBM copy
bandwidth, SURVEY 8(d) bytes = 16 + 24/A per update), `roofline_inner` (N = 1: the phi-only sweep + flux update the
inner iterations of a time step consist of, 12 flops per update against a DGEMM peak measured in this run),
`parity_check` (a small sharded k / time-march problem against the CPU oracle, run before anything is timed; a
mismatch aborts the run without a number), `cpu_baseline` and `e2e`.
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


def sweep_traffic(workload, world, cells):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the benchmarked sweep kernel, from the ncu capture
    committed under profiles/ for the CURRENT kernel version (profiles/sweep_traffic.json is keyed by it); None when
    there is no capture for this workload / rank count -- never a stale literal."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            t = json.load(f)
        with open(os.path.join(ROOT, "alpha_eigen_b200", "csrc", "ae_sweep.cu")) as f:
            import hashlib
            version = hashlib.sha1(f.read().encode()).hexdigest()[:12]
        e = t.get(f"{workload}_n{world}")
        if cells or e is None:
            return None, None
        src = e["source"] + ("" if e.get("kernel_sha1") == version else " [captured on an EARLIER version of ae_sweep.cu]")
        return float(e["dram_bytes_per_launch"]), src
    except Exception:
        return None, None


def measure_fp64_peak(torch, n=8192, reps=5):
    """cuBLAS DGEMM n^3, best of `reps` (CUDA events): the FP64 roofline denominator BASELINE.md asks to be measured
    on the box.  (tools/fp64_probe.cu measured 36.97 TFLOP/s DFMA issue, 37.14 DMMA on this pool: profiles/r02_fp64_peak.json.)"""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def parity_check(comm, rank, world):
    """Small sharded problems against the CPU oracle before anything is timed (the checker, not the thing measured):
    Kornreich-Parsons 20 cells/mfp S8 (one group: angle-sharded) k eigenvalue + 5 backward-Euler steps, and a 16-group
    synthetic S64 slab (sharded the way the benchmark problem is) source iteration + 3 steps.  Returns the dict that goes
    into the JSON line; `ok` false aborts the run."""
    from alpha_eigen_b200.device import SlabDevice
    from alpha_eigen_b200.synthetic import synthetic_problem
    from oracle import ae_oracle as orc
    from tests.kp import kp_problem, kp_psi0
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    p = kp_problem(9, 20, 8)
    t = p["tables"]
    mk = lambda prob, tab, **kw: SlabDevice.sharded(prob["Z"], prob["hx"], prob["mu"], prob["weight"], tab.mat, tab.total,
                                                    tab.scat, tab.chi, tab.nusf, tab.invspeed, comm=comm, **kw)
    dev = mk(p, t)
    rd = dev.power_iteration(p["phi0"])
    ro = orc.power_iteration(p["mu"], p["weight"], p["hx"], t, p["phi0"])
    k_err = abs(rd["k_new"] - ro["k_new"])
    phi_err = rel(rd["phi_new"], ro["phi_new"])
    iters = rd["status"] == 0 and list(rd["si_hist"]) == list(ro["si_hist"])
    devt = mk(p, t, dt=DT)
    psi = devt.psi_to_device(kp_psi0(p))
    st, outer, inner = devt.time_march(psi, 5, 1e-10)
    res = orc.time_march(p["mu"], p["weight"], p["hx"], t, np.zeros((p["Z"], 1)), kp_psi0(p), DT, 5)
    iters = iters and st == 0 and list(outer) == list(res["outer"]) and list(inner) == list(res["inner"])
    psi_err = rel(devt.psi_to_host(psi), res["psi_last"][:, devt.angles, :][..., devt.groups])
    # multigroup, streaming kernels (flags=1: no fused small-slab path), sharded like the benchmark problem
    q = synthetic_problem(1500, 64, 16)
    qt = orc.Tables(q["mat"], q["total"], q["scat"], q["chi"], q["nusf"], q["invspeed"])
    rs = np.random.RandomState(5)
    psi0 = rs.rand(q["Z"], q["A"], q["G"])
    s_ext = rs.rand(q["Z"], q["G"]) * 0.1
    devm = mk(q, qt, dt=DT, s_ext=s_ext, flags=1)
    pm = devm.psi_to_device(psi0)
    st, outer, inner = devm.time_march(pm, 3, 1e-10)
    rm = orc.time_march(q["mu"], q["weight"], q["hx"], qt, s_ext, psi0, DT, 3, want_psi_snap=False)
    iters = iters and st == 0 and list(outer) == list(rm["outer"]) and list(inner) == list(rm["inner"])
    mg_phi_err = rel(devm.cells_to_host(devm.phi), rm["phi_last"])
    mg_psi_err = rel(devm.psi_to_host(pm), rm["psi_last"][:, devm.angles, :][..., devm.groups])
    out = {"k_abs_err": k_err, "phi_rel_err": max(phi_err, mg_phi_err), "psi_rel_err": max(psi_err, mg_psi_err),
           "iters_match": bool(iters), "ranks": world, "sharding": [dev.shard_mode, devm.shard_mode],
           "problems": "K-P 20 cells/mfp S8 k + 5 steps; synthetic 16-group S64 1500 cells 3 steps (streaming kernels)"}
    out["ok"] = bool(iters and k_err < 1e-9 and out["phi_rel_err"] < 1e-10 and out["psi_rel_err"] < 1e-10)
    return out


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


# The reference arm / CPU baseline slice: the headline's angle and group counts on fewer cells (the literal reference
# cannot be instantiated at all at these sizes: one_direction.py:39 allocates (Z+1)^2 doubles per angle).
SLICE = dict(Z=20_000, A=256, G=70)


def host_cores():
    """Threads for the CPU legs: every core of the box.  torchrun exports OMP_NUM_THREADS=1 to its workers, which is
    about THEIR torch intra-op pools, not about this leg: it is ignored on purpose."""
    return os.cpu_count() or 1


class CpuSlice:
    """Oracle port (reference arithmetic, group.py:41-60 + time_unit.py:23) on the host cores: one `step` = one
    time-dependent sweep set of the slice (OpenMP over the groups)."""

    def __init__(self, Z=None, A=None, G=None):
        from alpha_eigen_b200.synthetic import synthetic_problem
        from oracle import ae_oracle as orc
        self.Z, self.A, self.G = Z or SLICE["Z"], A or SLICE["A"], G or SLICE["G"]
        p = synthetic_problem(self.Z, self.A, self.G)
        rs = np.random.RandomState(1)
        inv = p["invspeed"][p["mat"]]
        self.args = (p["mu"], p["weight"], p["hx"], rs.rand(self.Z, self.G), p["total"][p["mat"]] + 1 / DT * inv, inv / DT,
                     rs.rand(self.Z, self.A, self.G))
        self.cores = host_cores()
        os.environ["OMP_NUM_THREADS"] = str(self.cores)
        self.orc = orc
        self.updates = self.Z * self.A * self.G

    def step(self):
        self.orc.bench_sweep_sets(*self.args, 1)

    def describe(self, reps, seconds):
        return (f"{reps} time-dependent sweep sets of a synthetic {self.G}-group slab, Z={self.Z}, S{self.A} "
                f"({self.updates * reps:.3g} updates, psi {self.updates * 8 / 1e9:.2f} GB, {seconds:.1f} s); the literal "
                f"Python reference runs 1.6-1.75e5 updates/s on 1 core (BASELINE.md)")


def cpu_baseline(seconds=12.0):
    """`cpu_baseline` of the main arm: the same slice as `--impl reference`, time-bounded."""
    sl = CpuSlice()
    sl.step()                                                       # warm-up
    reps, t0 = 0, time.perf_counter()
    while reps < 2 or time.perf_counter() - t0 < seconds:
        sl.step()
        reps += 1
    el = time.perf_counter() - t0
    return {"value": sl.updates * reps / el, "unit": "cell*angle*group updates/s", "cores": min(sl.cores, sl.G),
            "kind": "port", "sample": sl.describe(reps, el)}


def run_reference(args, Z, A, G, desc):
    """--impl reference: the reference algorithm on the host cores (oracle port), W warm-up + K timed steps, each step
    one sweep set of the slice; ms_per_step is the measured time of such a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sl = CpuSlice()
    for _ in range(max(args.warmup, 1)):
        sl.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sl.step()
    el = time.perf_counter() - t0
    value = sl.updates * args.steps / el
    cb = {"value": value, "unit": "cell*angle*group updates/s", "cores": min(sl.cores, sl.G), "kind": "port",
          "sample": sl.describe(args.steps, el)}
    line = {"impl": "reference", "metric": "sweep cell*angle*group updates/s", "value": value,
            "unit": "updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": el * 1e3 / args.steps,
            "ms_per_step_what": f"one sweep set of the slice Z={sl.Z} (the headline problem has {Z} cells: x{Z / sl.Z:.0f})",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "Z": Z, "A": A, "G": G, "slice": {"Z": sl.Z, "A": sl.A, "G": sl.G}},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--skip-parity", action="store_true",
                    help="profiling runs only (ncu launch lists): do not launch the parity problems; the line says so")
    args = ap.parse_args()
    Z, A, G, desc = WORKLOADS[args.workload]
    if args.cells:
        Z = args.cells
        desc += f" [Z overridden to {Z}]"
    if args.impl == "reference":
        return run_reference(args, Z, A, G, desc)

    import ctypes as C

    import torch
    import torch.distributed as dist
    from alpha_eigen_b200 import _lib, distributed
    from alpha_eigen_b200.device import HostPipe, SlabDevice
    from alpha_eigen_b200.synthetic import synthetic_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = distributed.init() if world > 1 else None
    W = max(args.warmup, 3)

    # ---- correctness first: sharded k / march problems against the oracle, on every rank ------------------------
    parity = {"ok": True, "skipped": "--skip-parity (profiling run)"} if args.skip_parity else parity_check(comm, rank, world)
    if world > 1:
        ok = torch.tensor([1.0 if parity["ok"] else 0.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(ok.item() > 0.5)
    if not parity["ok"]:
        if rank == 0:
            print(json.dumps({"error": "parity check failed: no benchmark number", "parity_check": parity}), flush=True)
        if world > 1:
            distributed.shutdown()
        sys.exit(1)

    p = synthetic_problem(Z, A, G)
    force = os.environ.get("AE_BENCH_SHARD")              # "angle": the round-1 decomposition, for A/B runs
    if force == "angle" and comm is not None:
        from alpha_eigen_b200.device import shard_angles
        dev = SlabDevice(Z, p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                         p["invspeed"], dt=DT, angles=shard_angles(p["mu"], rank, world), comm=comm)
    else:
        dev = SlabDevice.sharded(Z, p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                                 p["invspeed"], comm=comm, dt=DT)
    psi = torch.empty(dev.Gl * dev.A * dev.Zp, dtype=torch.float64, device=dev.dev)
    psi.uniform_(0.5, 1.5, generator=torch.Generator(device=dev.dev).manual_seed(2021 + rank))
    # experiment knob: out-of-place psi update (needs a second psi buffer; the headline run is in place)
    psis = [psi, torch.empty_like(psi)] if os.environ.get("AE_BENCH_PSI_PINGPONG") == "1" else None
    dev.base.zero_()
    q = [dev.q0, dev.q1]
    q[0].uniform_(0.0, 1.0, generator=torch.Generator(device=dev.dev).manual_seed(7))
    updates_per_step = Z * A * G                        # whole job, all ranks
    local_updates = Z * dev.A * dev.Gl
    launches = 0
    angle_mode = comm is not None and dev.gl == 0

    group_mode = comm is not None and dev.gl > 0

    def step(i, ev=None):
        nonlocal launches
        if ev is not None:
            ev[0].record()
        pin, pout = (psis[i & 1], psis[(i + 1) & 1]) if psis is not None else (psi, psi)
        if group_mode:
            # group-sharded ranks: ONE call -- the groups are swept in two chunks and the first chunk's flux is assembled
            # and all-gathered on a second stream while the second chunk is swept; then the scatter source of the own groups
            dev.sweep_and_update(q[i & 1], pin, pout, q[(i + 1) & 1], 0.0, i + 1)
            launches += 6                                   # 2 sweeps, 2 assemblies, norm, scatter (+ 2 nccl, 16 copies)
            if ev is not None:
                ev[1].record(); ev[2].record(); ev[3].record()
            return
        _lib.check(dev.lib.ae_sweep_set(dev.slab, q[i & 1].data_ptr(), pin.data_ptr(), pout.data_ptr(),
                                        dev.partial.data_ptr(), dev.sync.data_ptr(), dev.stream))
        if ev is not None:
            ev[1].record()
        launches += 1
        if angle_mode:
            _lib.check(dev.lib.ae_exchange_partials(dev.slab, dev.work, dev.stream))   # local sum + reduce-scatter by rows
            if ev is not None:
                ev[2].record()
            dev.flux_update(dev.reduced, 1, q[(i + 1) & 1], 0.0, i + 1)             # 1/N of the cells + all-gather of q
            launches += 5                                   # reduce, (nccl), assemble, scatter, (nccl), norm
        else:
            if ev is not None:
                ev[2].record()
            # group-sharded ranks: assembly of the own groups, all-gather of the new flux, scatter source of the own groups
            dev.flux_update(dev.partial, dev.P, q[(i + 1) & 1], 0.0, i + 1)
            launches += 2 if comm is None else 4            # assemble, scatter (+ nccl, norm)
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
    if group_mode:
        phase = [float("nan")] * 3
    if world > 1:
        t = torch.tensor([ms_total] + phase, dtype=torch.float64, device=dev.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, phase = float(t[0]), [float(x) for x in t[1:]]
    sweep_ms = phase[0]
    sweep_ms_source = "CUDA events around the sweep launches of the timed steps"
    exchange_mode = None
    if group_mode and world > 1:
        peer = dev.lib.ae_comm_exchange_mode(dev.comm.handle) == 1          # what the library actually did
        split = peer and os.environ.get("AE_EXCHANGE_OVERLAP", "0") == "1" and G // world >= 3
        exchange_mode = (("copy engines over NVLink into IPC-shared peer buffers, flags by stream memory operations (no SM)"
                          if peer else "ncclAllGather through a staging buffer") +
                         ("; two group chunks, the first travels while the second is swept" if split else "; one chunk after the sweep"))
    if group_mode:
        # the step is ONE C call (ae_sweep_and_update) with no phase boundaries visible here: the sweep kernel is timed alone right after the
        # timed region (same launch, whole grid), the remainder of a step is exchange + update that the sweep did not hide
        fence()
        sev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for k in range(5):
            sev[k].record()
            _lib.check(dev.lib.ae_sweep_set(dev.slab, q[0].data_ptr(), psi.data_ptr(), psi.data_ptr(), dev.partial.data_ptr(),
                                            dev.sync.data_ptr(), dev.stream))
        sev[5].record()
        fence()
        sweep_ms = sev[0].elapsed_time(sev[5]) / 5
        if world > 1:
            t = torch.tensor([sweep_ms], dtype=torch.float64, device=dev.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sweep_ms = float(t[0])
        sweep_ms_source = "5 sweep-only launches timed right after the timed region (the step is one C call)"
        exchange_ms = None
        if world > 1:
            # the flux exchange alone (all-gather of the group rows of phi), ranks aligned by a barrier first
            dist.barrier()
            fence()
            xev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            xev[0].record()
            for k in range(5):
                _lib.check(dev.lib.ae_comm_gather_groups(dev.comm.handle, dev.phi.data_ptr(), G, dev.Zp, dev.stream))
            xev[1].record()
            fence()
            t = torch.tensor([xev[0].elapsed_time(xev[1]) / 5], dtype=torch.float64, device=dev.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            exchange_ms = float(t[0])
    ms_per_step = ms_total / args.steps
    value = updates_per_step / (ms_per_step * 1e-3)
    gpu_launches = launches
    assert bool(torch.isfinite(dev.phi).all()), "non-finite scalar flux"

    # ---- the inner-iteration variant (N = 1): phi-only sweep set + flux update, what a time step spends its time in ----
    inner = None
    if comm is None:
        Pw = C.c_int32(0)

        def inner_step(i, ev=None):
            if ev is not None:
                ev[0].record()
            _lib.check(dev.lib.ae_sweep_set_phi(dev.slab, q[i & 1].data_ptr(), dev.partial.data_ptr(), dev.sync.data_ptr(),
                                                C.byref(Pw), dev.stream))
            if ev is not None:
                ev[1].record()
            dev.flux_update(dev.partial, Pw.value, q[(i + 1) & 1], 0.0, i + 1)
            if ev is not None:
                ev[2].record()
        for i in range(3):
            inner_step(i)
        fence()
        ievs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        for i in range(args.steps):
            inner_step(i, ievs[i])
        fence()
        i_sweep = float(np.mean([e[0].elapsed_time(e[1]) for e in ievs]))
        i_update = float(np.mean([e[1].elapsed_time(e[2]) for e in ievs]))
        fp64_peak = measure_fp64_peak(torch)
        flops = 12.0 * local_updates                          # 12 FP64 flops per update, one a division
        inner = {"bound": "fp64", "achieved": flops / (i_sweep * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                 "frac": flops / (i_sweep * 1e-3) / 1e12 / fp64_peak, "traffic": None,
                 "kernel": "ae::phi_sweep_kernel (4 angles per warp, 32-angle blocks, uniform-tile constants)",
                 "kernel_ms": i_sweep, "updates_per_s": local_updates / (i_sweep * 1e-3),
                 "flops_per_update": 12.0, "algorithmic_flops": flops,
                 "algorithmic_bytes": 24.0 / dev.A * local_updates,
                 "hbm_GBps_algorithmic": 24.0 / dev.A * local_updates / (i_sweep * 1e-3) / 1e9,
                 "flux_update_ms": i_update, "iteration_ms": i_sweep + i_update,
                 "peak_source": "cuBLAS DGEMM 8192^3, best of 5, measured in this run (tools/fp64_probe.cu: 36.97 TFLOP/s "
                                "DFMA issue, 37.14 DMMA on this pool)",
                 "note": "the kernel does fewer flops than the 12 per update the reference spends: in tiles whose cells "
                         "share one sigT the reciprocal and the coefficient products are per-row constants; ncu "
                         "(profiles/r02_phi_sweep_c3_ncu.md): no pipe saturated, bound by the latency of the fold / scan chain"}

    # ---- end to end through host buffers: the source comes from pinned host memory, the flux goes back ---------
    # Several GPUs: host I/O is sharded like the flux update.  Group-sharded ranks upload the source rows of THEIR groups
    # (all a rank ever reads) and read back the flux of their groups: 1/N of the bytes each, no exchange.  Angle-sharded
    # ranks upload their cell range, all-gather the ranges over NVLink and read back their range of the flux.
    pipe = HostPipe(dev)
    hz, hg = pipe.shard_shape()
    q_host = torch.rand(max(hz, 1), hg, dtype=torch.float64)[:hz].pin_memory()
    phi_host = torch.empty(hz, hg, dtype=torch.float64).pin_memory()
    e_steps = max(3, min(args.steps, 32))       # the pipe's fill (first upload) and drain (last download) are inside the timed region

    # Every step still uploads its own source and downloads its own flux; the copies run on the copy engines
    # (HostPipe: two staging slots, own streams) so that the upload of step i+1 and the download of step i-1
    # overlap the kernels of step i.  AE_BENCH_E2E_SERIAL=1 times the unpipelined sequence instead.
    serial = os.environ.get("AE_BENCH_E2E_SERIAL") == "1"

    def e2e_run(n):
        if serial:
            for i in range(n):
                pipe.upload(i, q_host)
                pipe.feed(i, q[i & 1])
                step(i)
                pipe.drain(i, dev.phi, phi_host)
                pipe.wait()
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
        bytes_per_update = 16.0 + 24.0 / dev.A                         # psi read + write, hs/q/cdt over the angles
        alg_bytes = bytes_per_update * local_updates
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
        traffic, traffic_src = sweep_traffic(args.workload, world, args.cells)
        sharding = {"single": "one GPU", "group": f"whole groups ({dev.Gl} of {G} on rank 0) x all angles: no flux reduction, one "
                    "all-gather of the new flux per step", "angle": f"angles of every group dealt over {world} ranks: flux "
                    "reduce-scatter + all-gather of the new source per step"}[dev.shard_mode]
        line = {
            "metric": "sweep cell*angle*group updates/s", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "Z": Z, "A": A, "G": G, "dt": DT, "rows_per_gpu": dev.A * dev.Gl,
                       "sharding": sharding,
                       "step": "psi-streaming sweep set (in place) + flux exchange (N>1) + flux-update kernels",
                       "l2": "inputs (psi %.1f GB per GPU) are larger than L2" % (psi.numel() * 8 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes,
                         "kernel": "ae::sweep_kernel<16,true,true> (8 for fewer than 16 angles per direction)",
                         "kernel_ms": sweep_ms, "kernel_ms_source": sweep_ms_source, "bytes_per_update": bytes_per_update,
                         "peak_source": peak_src,
                         "sweep_share_of_step": sweep_ms / ms_per_step},
            "roofline_inner": inner,
            "breakdown_ms": ({"sweep_alone": sweep_ms, "exposed_exchange_and_update": ms_per_step - sweep_ms,
                              "exchange_alone": exchange_ms, "exchange": exchange_mode,
                              "what": "group-sharded step: the whole-range sweep kernel timed alone; the rest of the step is what "
                                      "the sweep did not hide (flux assembly, the exposed part of the flux exchange, norm, scatter "
                                      "source, and what splitting the sweep in two launches costs)"}
                             if group_mode else
                             {"sweep": phase[0], "exchange_reduce_scatter": phase[1], "flux_update_and_allgather": phase[2]}),
            "parity_check": parity,
            "e2e": {"value": updates_per_step / (e_ms * 1e-3), "unit": "updates/s", "ms_per_step": e_ms,
                    "h2d_bytes_per_step": Z * G * 8, "d2h_bytes_per_step": Z * G * 8,
                    "steps": e_steps, "pipelined": not serial,
                    "what": "whole-job bytes; every step every rank copies its share of the source (its groups, or its cell "
                            "range) from pinned host memory, runs the step on resident psi, and reads its share of the flux "
                            "back to pinned host memory; copies run on their own streams, double-buffered, so the upload of "
                            "step i+1 and the download of step i-1 overlap the kernels of step i"},
            "gpu_launches": gpu_launches, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        distributed.shutdown()


if __name__ == "__main__":
    main()
