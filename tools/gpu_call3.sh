#!/bin/bash
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite exit $? in ${SECONDS}s"; tail -5 gpurun_out/gpu_tests.log
for n in 2 3 4; do
  AE_ENSEMBLE_CTAS_PER_SM=$n timeout 600 python - > gpurun_out/c5_minb$n.log 2>&1 <<PY
import json, sys
sys.path.insert(0, '.')
import bench_eig
bench_eig.case_c5_batched(lambda d: print(json.dumps(d)), E=128)
PY
  echo "minb=$n: $(tail -1 gpurun_out/c5_minb$n.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("members_in_flight","wall_s","setup_s","march_s","dmd_s","all_converged")})' 2>&1 | tail -1)"
done
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_pipe.json 2> gpurun_out/bench_c4_pipe.err
echo "bench exit $?"
AE_BENCH_E2E_SERIAL=1 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_serial.json 2> gpurun_out/bench_c4_serial.err
python - <<'PY'
import json
for f in ("pipe", "serial"):
    try:
        d = json.loads(open(f"gpurun_out/bench_c4_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step %.2f sweep %.2f frac %.3f e2e %.2f ms (%s)" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["clocks"]))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/bench_c4_{f}.err").read()[-2000:])
PY
