#!/bin/bash
# 2-GPU development check: both decompositions against the oracle with the peer-copy flux exchange and with the NCCL one,
# the two-chunk step against the plain calls, bench A/B of the three exchange variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
AE_COMM_PEER_COPY=1 timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/n2_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"
grep "mgpu_check ok\|Error\|error\|exchange" gpurun_out/n2_mgpu_check.log | cut -c1-260 | head -6
timeout 300 $TR --master-port 29514 tests/mgpu_check.py > gpurun_out/n2_mgpu_check_nccl.log 2>&1; echo "mgpu_check (NCCL exchange) rc=$?"
grep "Error\|error\|exchange" gpurun_out/n2_mgpu_check_nccl.log | cut -c1-260 | head -4
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline"
AE_COMM_PEER_COPY=1 AE_EXCHANGE_OVERLAP=1 timeout 400 $TR --master-port 29512 $B > gpurun_out/n2_bench_peer_split.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"
AE_COMM_PEER_COPY=1 timeout 400 $TR --master-port 29513 $B > gpurun_out/n2_bench_peer_one.json 2> /dev/null; echo "bench (one chunk) rc=$?"
timeout 400 $TR --master-port 29515 $B > gpurun_out/n2_bench_nccl.json 2> /dev/null; echo "bench (NCCL) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n2_bench_peer_split.json", "gpurun_out/n2_bench_peer_one.json", "gpurun_out/n2_bench_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        b = d["breakdown_ms"]
        print(f, d["ms_per_step"], b["sweep_alone"], b["exposed_exchange_and_update"], b["exchange"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -c 600 gpurun_out/n2_bench.err | grep -v Warn | tail -5
