#!/bin/bash
# Round-end verification on one GPU; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; echo "gpu suite exit $? (${SECONDS}s): $(tail -1 gpurun_out/final_gpu_tests.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?: $(tail -1 gpurun_out/final_smoke.log)"
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 900 python bench_eig.py --out gpurun_out/final_eig_bench.jsonl > gpurun_out/final_eig_bench.log 2>&1; echo "eig bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("bench: ms/step %.2f value %.3e sweep %.2f frac %.3f e2e %.2f ms (%.3e) launches %d clocks %s" % (d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]))
print("cpu_baseline:", d.get("cpu_baseline"))
r = json.loads(open("gpurun_out/final_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm: %.3e %s cores %s" % (r["value"], r["unit"], r["cpu_baseline"]["cores"]))
for l in open("gpurun_out/final_eig_bench.jsonl"):
    e = json.loads(l)
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in e.items() if k in ("case", "cells_per_mfp", "num_angles", "steps", "gpu_wall_s", "gpu_march_s", "gpu_dmd_s", "gpu_dmd_tsqr_s", "gpu_dmd_gram_s", "k_minus_golden", "alpha0", "alpha0_rel_diff", "iterations_match", "wall_s", "march_s", "dmd_s", "members", "updates_per_s")})
PY
