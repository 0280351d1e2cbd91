#!/bin/bash
# round 2, final single-GPU measurement batch: tests, the contract bench, launch lists, ncu captures, the eigenvalue bench.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/f_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/f_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2>&1; echo "ref arm rc=$?"
# launch list of the same bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu1.log 2>&1
# the benchmarked kernel, full set
python tools/sweep_bench.py --shape c4 --reps 3 --tag t > gpurun_out/f_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o gpurun_out/r02_sweep_c4_rw \
    python tools/sweep_bench.py --shape c4 --reps 3 --tag t > gpurun_out/f_ncu2.log 2>&1
python tools/sweep_bench.py --shape c3 --tag final > gpurun_out/r02_sweep_c3_final.jsonl 2>&1
python tools/sweep_bench.py --shape c4 --reps 5 --tag final > gpurun_out/r02_sweep_c4_final.jsonl 2>&1
python bench_eig.py --out gpurun_out/r02_eig_bench.jsonl > gpurun_out/f_eig.log 2>&1; echo "eig bench rc=$?"
python bench_eig.py --only-c3 --c3-steps 10 --out gpurun_out/g.jsonl > gpurun_out/f_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r02_launches_c3_march.csv \
    python bench_eig.py --only-c3 --c3-steps 10 --out gpurun_out/g.jsonl > gpurun_out/f_ncu3.log 2>&1
tail -3 gpurun_out/r02_eig_bench.jsonl | cut -c1-300
