#!/bin/bash
# round 2, the 8-GPU call: correctness at N=8 in both decompositions (and the overlapped step against the plain calls),
# the C4 bench with and without the exchange overlap.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/n8_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; grep "mgpu_check ok\|Error" gpurun_out/n8_mgpu_check.log | cut -c1-300 | head -3
timeout 500 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_bench_overlap.json 2> gpurun_out/n8_bench_overlap.err; echo "bench overlap rc=$?"
AE_EXCHANGE_OVERLAP=0 timeout 500 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_bench_no_overlap.json 2> /dev/null; echo "bench no-overlap rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n8_bench_overlap.json", "gpurun_out/n8_bench_no_overlap.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
