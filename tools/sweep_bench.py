#!/usr/bin/env python
"""Kernel A/B probe: times the sweep-set variants (psi read+write, psi read-only, phi-only) on a benchmark shape,
with and without the uniform-tile records, CUDA events on the launching stream.  One JSON line per variant."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c3", choices=["c3", "c4", "c4s"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    from alpha_eigen_b200.device import SlabDevice
    from alpha_eigen_b200.synthetic import synthetic_problem
    Z, A, G = {"c3": (100_000, 64, 70), "c4": (1_000_000, 256, 70), "c4s": (250_000, 256, 70)}[args.shape]
    p = synthetic_problem(Z, A, G)
    dev = SlabDevice(Z, p["hx"], p["mu"], p["weight"], p["mat"], p["total"], p["scat"], p["chi"], p["nusf"],
                     p["invspeed"], dt=0.1)
    psi = torch.empty(dev.G * dev.A * dev.Zp, dtype=torch.float64, device=dev.dev)
    psi.uniform_(0.5, 1.5)
    q = dev.q0
    q.uniform_(0.0, 1.0)
    info_ptr = dev.t_tinfo.data_ptr()
    upd = Z * A * G
    for use_info in (True, False):
        dev.slab.tile_info = info_ptr if use_info else None
        for name, pin, pout in (("rw", psi, psi), ("ro", psi, None), ("phi", None, None)):
            ptr = lambda t: t.data_ptr() if t is not None else None
            import ctypes as C
            Pw = C.c_int32(0)
            def run():
                if name == "phi":       # the kernel the solver loops use for phi-only iterations
                    rc = dev.lib.ae_sweep_set_phi(dev.slab, q.data_ptr(), dev.partial.data_ptr(), dev.sync.data_ptr(),
                                                  C.byref(Pw), dev.stream)
                else:
                    rc = dev.lib.ae_sweep_set(dev.slab, q.data_ptr(), ptr(pin), ptr(pout), dev.partial.data_ptr(),
                                              dev.sync.data_ptr(), dev.stream)
                assert rc == 0, rc
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            ev[0].record()
            for i in range(args.reps):
                run()
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
            med = ms[len(ms) // 2]
            bpu = {"rw": 16 + 24 / A, "ro": 8 + 24 / A, "phi": 24 / A}[name]
            print(json.dumps({"tag": args.tag, "shape": args.shape, "variant": name, "uniform_tiles": use_info,
                              "ms_median": med, "ms_min": ms[0], "updates_per_s": upd / (med * 1e-3),
                              "alg_GBps": bpu * upd / (med * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
