#!/bin/bash
# round 2, the 8-GPU call: correctness at N=8 in both decompositions, the C4 bench (group-sharded, and the round-1 angle
# decomposition for comparison), BASELINE configs[4] on the whole box.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 tests/mgpu_check.py > gpurun_out/n8_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; grep "mgpu_check ok\|Error" gpurun_out/n8_mgpu_check.log | head -3
$TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_bench_group.json 2> gpurun_out/n8_bench_group.err; echo "bench group rc=$?"
AE_BENCH_SHARD=angle $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_bench_angle.json 2> gpurun_out/n8_bench_angle.err; echo "bench angle rc=$?"
$TR --master-port 29514 tools/c5_multi.py --members 1024 --row-sharded 8 > gpurun_out/n8_c5.jsonl 2> gpurun_out/n8_c5.err; echo "c5 rc=$?"
tail -c 400 gpurun_out/n8_c5.jsonl
