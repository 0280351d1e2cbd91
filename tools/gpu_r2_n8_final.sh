#!/bin/bash
# round 2, last 8-GPU call: the driver's own bench command at N=8 and BASELINE configs[4] on the whole box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench_final.json 2> gpurun_out/n8_bench_final.err; echo "bench rc=$?"
timeout 500 $TR --master-port 29531 tools/c5_multi.py --row-sharded 2 > gpurun_out/n8_c5.jsonl 2> gpurun_out/n8_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n8_bench_final.json").read().strip().splitlines()[-1])
b = d["breakdown_ms"]
print(d["value"], d["ms_per_step"], b["sweep_alone"], b["exposed_exchange_and_update"], b["exchange_alone"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["steps"], "parity", d["parity_check"]["ok"])
PY
grep '"case"' gpurun_out/n8_c5.jsonl | cut -c1-420
