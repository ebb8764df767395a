#!/bin/bash
python tools/prep_bench.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:remap_kernel -s 3 -c 1 -f -o gpurun_out/r01d_remap python tools/prep_bench.py > gpurun_out/ncu_remap.log 2>&1
python tools/ncu_summary.py gpurun_out/r01d_remap.ncu-rep > gpurun_out/r01d_remap_ncu_summary.txt 2>&1
