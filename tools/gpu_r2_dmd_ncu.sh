#!/bin/bash
# ncu --set full of the round's final DMD kernels: gram_kernel on the C3 window, dmd_small_kernel on a real member window
mkdir -p gpurun_out
python tools/dmd_probe.py > gpurun_out/y_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 9 -c 1 -o gpurun_out/r02_gram_c3_final -f \
    python tools/dmd_probe.py > gpurun_out/y_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmd_small_kernel -s 1 -c 1 -o gpurun_out/r02_dmd_small_final -f \
    python tools/dmd_probe.py > gpurun_out/y_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
