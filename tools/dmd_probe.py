"""Time ae_gram and ae_tsqr alone (CUDA events) on the two benchmark window shapes; one JSON line each.  GPU only."""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from alpha_eigen_b200 import _lib  # noqa: E402

lib = _lib.load()
n = 116
st = torch.cuda.current_stream().cuda_stream
for name, M in (("c5_member_psi", 393216), ("c3_phi", 7_006_720)):
    snap = torch.rand(n * M, dtype=torch.float64, device="cuda")
    out = torch.empty(n * n, dtype=torch.float64, device="cuda")
    for kind, fn in (("gram", lib.ae_gram), ("tsqr", lib.ae_tsqr)):
        for _ in range(2):
            _lib.check(fn(snap.data_ptr(), M, M, n, out.data_ptr(), st))
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(snap.data_ptr(), M, M, n, out.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        best = min(ms)
        # Gram: M n (n+1) flops (SYRK form); TSQR (Householder, R only): 2 M n^2
        flops = M * n * (n + 1) if kind == "gram" else 2.0 * M * n * n
        print(json.dumps({"shape": name, "M": M, "n": n, "kernel": kind, "ms": best, "GBps": n * M * 8 / (best * 1e-3) / 1e9,
                          "TFLOPs_algorithmic": flops / (best * 1e-3) / 1e12}), flush=True)
    del snap

# ---- the small dense part on a real window: one ensemble member's march, then ae_dmd_from_r through the block kernel
# and through cuSOLVER, and 128 windows at once (wall time: the calls end with a stream synchronisation)
import os  # noqa: E402
import time  # noqa: E402

from alpha_eigen_b200.ensemble import solve_member, synthetic_ensemble  # noqa: E402
from alpha_eigen_b200.time_unit import dmd_window  # noqa: E402

p = synthetic_ensemble(1024, 2048, 16, 12)[0]
res = solve_member(p, 0.1, 200, keep_snapshots=True)
dev, snap = res["dev"], res["snap"]
rows = snap.numel() // 200
first, K = dmd_window(200)
R = dev.tsqr(snap, rows, first, K + 1)


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


os.environ.pop("AE_DMD_CUSOLVER", None)
ms_k = wall(lambda: dev.dmd_from_r(R, K, 0.1))
sweeps = lib.ae_dmd_last_jacobi_sweeps()
stack = R.repeat(128)
ms_b = wall(lambda: dev.dmd_from_r_batched(stack, 128, K, 0.1), reps=3)
os.environ["AE_DMD_CUSOLVER"] = "1"
ms_l = wall(lambda: dev.dmd_from_r(R, K, 0.1))
os.environ.pop("AE_DMD_CUSOLVER", None)
print(json.dumps({"shape": "c5_member_window", "K": K, "rank": int(res["rank"]), "dmd_from_r_block_kernel_ms": ms_k,
                  "jacobi_sweeps": int(sweeps), "dmd_from_r_batched_128_windows_ms": ms_b, "dmd_from_r_cusolver_ms": ms_l}), flush=True)
