#!/bin/bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite exit $? in ${SECONDS}s"; tail -8 gpurun_out/gpu_tests.log
timeout 300 python tools/dmd_split.py > gpurun_out/dmd_split2.log 2>&1; echo "dmd split exit $?"; tail -4 gpurun_out/dmd_split2.log
timeout 900 python bench_eig.py --only-c5 --out gpurun_out/c5_bench2.jsonl > gpurun_out/c5_bench2.log 2>&1
echo "c5 bench exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/c5_bench2.jsonl'):
    d = json.loads(l)
    print({k: d.get(k) for k in ("case", "snapshots", "members", "members_in_flight", "wall_s", "setup_s", "march_s", "dmd_s", "all_converged", "ranks", "alpha0_first_last")})
PY
