"""Time ae_tsqr alone on random tall matrices [n][M] (snapshot-row layout); GPU only."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from alpha_eigen_b200 import _lib  # noqa: E402

lib = _lib.load()
n = 116
for M in (393216, 1_000_000, 7_006_720):
    snap = torch.rand(n * M, dtype=torch.float64, device="cuda")
    r = torch.empty(n * n, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(lib.ae_tsqr(snap.data_ptr(), M, M, n, r.data_ptr(), st))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"M={M} n={n} rep {rep}: {1e3 * (t1 - t0):.2f} ms  ({n * M * 8 / (t1 - t0) / 1e9:.0f} GB/s)", flush=True)
    del snap
