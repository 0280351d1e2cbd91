#!/bin/bash
# Batched-ensemble checks on one GPU: its tests, the whole GPU suite (timed), the configs[4] bench.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_public_api.py -x -q -k "ensemble" > gpurun_out/ens_tests.log 2>&1
echo "ensemble tests exit $?"; tail -15 gpurun_out/ens_tests.log
SECONDS=0
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite exit $? in ${SECONDS}s"; tail -5 gpurun_out/gpu_tests.log
timeout 900 python bench_eig.py --only-c5 --out gpurun_out/c5_bench.jsonl > gpurun_out/c5_bench.log 2>&1
echo "c5 bench exit $?"; tail -5 gpurun_out/c5_bench.log
