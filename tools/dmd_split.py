"""Time the two halves of a member DMD (ae_tsqr, ae_dmd_from_r) on one configs[4] member; GPU only."""
import json
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from alpha_eigen_b200.ensemble import solve_members_batched, synthetic_ensemble  # noqa: E402
from alpha_eigen_b200.time_unit import dmd_window  # noqa: E402

members = synthetic_ensemble(1024, 2048, 16, 12)[::512][:2]
res = solve_members_batched(members, 0.1, 200, keep_snapshots=True)
r = res[0]
dev, snap = r["dev"], r["snap"]
rows = dev.G * dev.A * dev.Zp
first, K = dmd_window(200)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    R = dev.tsqr(snap, rows, first, K + 1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    al, sig, rk = dev.dmd_from_r(R, K, 0.1)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"rows": rows, "cols": K + 1, "tsqr_ms": (t1 - t0) * 1e3, "dmd_from_r_ms": (t2 - t1) * 1e3, "rank": int(rk),
                      "alpha0": float(max(al.real))}))
