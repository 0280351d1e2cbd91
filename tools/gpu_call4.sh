#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_public_api.py -x -q -k "cli" > gpurun_out/cli_tests.log 2>&1; echo "cli tests exit $?"; tail -3 gpurun_out/cli_tests.log
cat > /tmp/ens_small.py <<'PY'
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from alpha_eigen_b200.ensemble import solve_members_batched, synthetic_ensemble, member_dmd
from alpha_eigen_b200.time_unit import dmd_window
mode = sys.argv[1]
members = synthetic_ensemble(1024, 2048, 16, 12)[::32][:24]
if mode == "march":
    res = solve_members_batched(members, 0.1, 12, snapshots="phi", dmd_threads=1)
    print("ok", [r["status"] for r in res][:4])
else:
    res = solve_members_batched(members[:2], 0.1, 200, keep_snapshots=True, dmd_threads=1)
    r = res[0]; dev = r["dev"]; snap = r["snap"]; rows = dev.G * dev.A * dev.Zp
    first, K = dmd_window(200)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        R = dev.tsqr(snap, rows, first, K + 1)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        al, sig, rk = dev.dmd_from_r(R, K, 0.1)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(json.dumps({"tsqr_ms": (t1 - t0) * 1e3, "dmd_from_r_ms": (t2 - t1) * 1e3, "rank": int(rk)}))
PY
timeout 300 python /tmp/ens_small.py dmd > gpurun_out/dmd_split.log 2>&1; echo "dmd split exit $?"; tail -3 gpurun_out/dmd_split.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dmd_launches.csv python /tmp/ens_small.py dmd > gpurun_out/dmd_ncu.log 2>&1; echo "dmd launch list exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/dmd_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4][:70]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print("kernels", len(rows), "total us", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]: print("%6d %10.1f us  %s" % (v[0], v[1], k))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensemble_time_march -c 1 -o gpurun_out/ens_march_full python /tmp/ens_small.py march > gpurun_out/ens_ncu.log 2>&1; echo "ens ncu exit $?"; tail -3 gpurun_out/ens_ncu.log
ls -la gpurun_out/*.ncu-rep
