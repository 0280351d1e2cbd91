#!/bin/bash
# A/B of the sweep kernel's gang size and L2 hints on C4 (one GPU); results under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "sweep" > gpurun_out/sweep_tests.log 2>&1
echo "sweep tests exit $?" | tee -a gpurun_out/sweep_tests.log
tail -3 gpurun_out/sweep_tests.log
: > gpurun_out/sweep_ab.jsonl
for cfg in "1 0" "1 1" "auto 1" "auto 0" "8 1" "2 1"; do
  set -- $cfg
  echo "# AE_SWEEP_GANG=$1 AE_SWEEP_L2_HINTS=$2" >> gpurun_out/sweep_ab.jsonl
  AE_SWEEP_GANG=$1 AE_SWEEP_L2_HINTS=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline >> gpurun_out/sweep_ab.jsonl 2> gpurun_out/sweep_ab_err.log
done
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:sweep_kernel --launch-skip 3 --launch-count 1 --csv --log-file gpurun_out/c4_sweep_dram_gang.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/sweep_ab.jsonl'):
    if l.startswith('#'): print(l.strip()); continue
    try:
        d=json.loads(l); r=d['roofline']
        print('  ms/step %.2f sweep %.2f frac %.3f clocks %s e2e_ms %.1f'%(d['ms_per_step'], r['kernel_ms'], r['frac'], d['clocks'], d['e2e']['ms_per_step']))
    except Exception as e: print('  bad line', e, l[:200])
PY
grep -v "^==" gpurun_out/c4_sweep_dram_gang.csv | tail -5
