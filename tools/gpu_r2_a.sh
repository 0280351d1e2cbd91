#!/bin/bash
# round 2, call A: new tests first, then the whole GPU suite, then kernel A/B numbers
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "uniform_tile or fixed_flux or streaming_fixed" > gpurun_out/a_newtests.log 2>&1; echo "new tests rc=$?" 
tail -15 gpurun_out/a_newtests.log
python -m pytest tests -x -q -m gpu > gpurun_out/a_gputests.log 2>&1; echo "gpu suite rc=$?"
tail -15 gpurun_out/a_gputests.log
python tools/sweep_bench.py --shape c3 --tag r2a > gpurun_out/a_sweep_c3.jsonl 2>&1; cat gpurun_out/a_sweep_c3.jsonl
python tools/sweep_bench.py --shape c4 --reps 5 --tag r2a > gpurun_out/a_sweep_c4.jsonl 2>&1; cat gpurun_out/a_sweep_c4.jsonl
