#!/bin/bash
# round 2, late single-GPU batch: the whole GPU suite, the DMD kernel probe, the bench launch list, the sweep's DRAM bytes
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/s_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/s_gputests.log
python tools/dmd_probe.py > gpurun_out/r02_dmd_probe_v2.jsonl 2> gpurun_out/s_dmd.err; echo "dmd probe rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-parity > gpurun_out/s_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-parity > gpurun_out/s_ncu1.log 2>&1
python tools/sweep_bench.py --shape c4 --reps 3 --tag t > gpurun_out/s_plain2.log 2>&1 && \
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/r02_c4_sweep_dram_bytes.csv python tools/sweep_bench.py --shape c4 --reps 3 --tag t > gpurun_out/s_ncu2.log 2>&1
grep -c sweep_kernel gpurun_out/r02_launches_c4.csv; grep "dram__bytes" gpurun_out/r02_c4_sweep_dram_bytes.csv | cut -d, -f5,13,15
