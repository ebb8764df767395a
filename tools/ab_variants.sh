#!/bin/bash
# A/B of prebuilt library variants (_variants/liborbx_<name>.so, built here with different -D switches): each is copied over the
# in-tree liborbx.so on the GPU box, the short bench runs, the base library is restored at the end.
SO=wut_cuda_orb_slam3_b200/liborbx.so
cp $SO /tmp/liborbx_keep.so
for v in "$@"; do
  cp _variants/liborbx_$v.so $SO
  python bench.py --steps 10 --warmup 3 --no-knn2 --no-other --no-cpu ${AB_EXTRA:---no-cfg4} 2>/dev/null | python tools/bench_brief.py $v
done
cp /tmp/liborbx_keep.so $SO
