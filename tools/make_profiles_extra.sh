#!/bin/bash
# Run on the GPU box (gpurun): `ncu --set full` captures of the kernels that tools/make_profiles.sh does not cover — the octree
# (second-largest stage), orientation + descriptors and the tiled pyramid resize — for the round's evidence and the next round's
# work list.  Each ncu command is preceded by the same command without ncu; numbers printed under ncu are never used.
set -u
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --batch 64 --no-knn2 --no-cpu --no-other"
python bench.py $ARGS > $OUT/plain.log 2>&1 || { tail -5 $OUT/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:octree_kernel -s 1 -c 1 -f -o $OUT/${TAG}_octree python bench.py $ARGS > $OUT/ncu_o.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orient_describe -s 1 -c 1 -f -o $OUT/${TAG}_orient_describe python bench.py $ARGS > $OUT/ncu_od.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pyr_resize_tiled -s 7 -c 1 -f -o $OUT/${TAG}_pyr_resize python bench.py $ARGS > $OUT/ncu_pr.log 2>&1
tail -n 2 $OUT/ncu_o.log $OUT/ncu_od.log $OUT/ncu_pr.log
