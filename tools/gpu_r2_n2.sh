#!/bin/bash
# 2-GPU development check: both decompositions against the oracle, the overlapped step against the plain calls, bench A/B
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/n2_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"
grep "mgpu_check ok\|Error\|error" gpurun_out/n2_mgpu_check.log | cut -c1-220 | head -5
timeout 400 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"
AE_EXCHANGE_OVERLAP=1 timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench_overlap.json 2> /dev/null; echo "bench (overlap forced) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n2_bench.json", "gpurun_out/n2_bench_overlap.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["breakdown_ms"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/n2_bench.err | grep -v Warn | tail -5
