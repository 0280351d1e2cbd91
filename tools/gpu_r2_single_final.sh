#!/bin/bash
# round 2, last single-GPU batch: whole GPU suite, the driver's bench command, launch list, eigenvalue bench, DMD probe
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/z_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/z_gputests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-parity > gpurun_out/z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-parity > gpurun_out/z_ncu1.log 2>&1
python tools/dmd_probe.py > gpurun_out/r02_dmd_probe.jsonl 2> gpurun_out/z_dmd.err; echo "dmd probe rc=$?"
python bench_eig.py --out gpurun_out/r02_eig_bench.jsonl > gpurun_out/z_eig.log 2>&1; echo "eig bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_c4_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic_source"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["steps"], "inner", (d.get("roofline_inner") or {}).get("frac"))
PY
grep -c sweep_kernel gpurun_out/r02_launches_c4.csv
grep '"case": "c[35]' gpurun_out/r02_eig_bench.jsonl | cut -c1-330
