#!/bin/bash
# round 2, the 8-GPU call: correctness at N=8 in both decompositions (peer-copy flux exchange; the two-chunk step against the
# plain calls), the C4 bench with the peer-copy exchange and with the NCCL all-gather.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
AE_COMM_PEER_COPY=1 timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/n8_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; grep "mgpu_check ok\|Error\|exchange" gpurun_out/n8_mgpu_check.log | cut -c1-300 | head -4
B="bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline"
AE_COMM_PEER_COPY=1 timeout 500 $TR --master-port 29512 $B > gpurun_out/n8_bench_peer.json 2> gpurun_out/n8_bench_peer.err; echo "bench peer rc=$?"
timeout 500 $TR --master-port 29513 $B > gpurun_out/n8_bench_nccl.json 2> /dev/null; echo "bench nccl rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/n8_bench_peer.json", "gpurun_out/n8_bench_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        b = d["breakdown_ms"]
        print(f, d["value"], d["ms_per_step"], b["sweep_alone"], b["exposed_exchange_and_update"], b["exchange_alone"], b["exchange"][:40], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_check"]["ok"], "frac", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
