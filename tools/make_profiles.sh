#!/bin/bash
# Run on the GPU box (gpurun): final bench + ncu evidence for profiles/.  Every ncu command is preceded by the same
# command without ncu (B200_PROFILING.md).
set -u
OUT=gpurun_out
ARGS="--steps 1 --warmup 1 --batch 64 --no-knn2 --no-cpu"
python bench.py --steps 10 --warmup 3 > $OUT/final_bench.json 2> $OUT/final_bench.err
python bench.py $ARGS > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/final_launches.csv python bench.py $ARGS > $OUT/ncu_l.log 2>&1
python bench.py $ARGS > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fast_cells -s 1 -c 1 -o $OUT/final_fast python bench.py $ARGS > $OUT/ncu_f.log 2>&1
KARGS="--steps 1 --warmup 1 --batch 32 --no-cpu --knn-ndb 1000000 --knn-reps 1"
python bench.py $KARGS > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn2_kernel -s 1 -c 1 -o $OUT/final_knn python bench.py $KARGS > $OUT/ncu_k.log 2>&1
python tools/latency.py > $OUT/final_latency.txt 2>&1
tail -n 2 $OUT/ncu_f.log; tail -n 2 $OUT/ncu_k.log
