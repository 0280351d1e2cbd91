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
