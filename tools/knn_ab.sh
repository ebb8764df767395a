#!/bin/bash
# same-box A/B of matcher variants (_variants/liborbx_<name>.so) with the tensor-core path on: 100 k x 10 M, verified
SO=wut_cuda_orb_slam3_b200/liborbx.so
cp $SO /tmp/liborbx_keep.so
for v in "$@"; do
  cp _variants/liborbx_$v.so $SO
  ORBX_KNN_IMMA=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-other --no-cfg4 --knn-reps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); k = d['knn2']; print('$v', '%.4e' % k['value'], k['verified'], round(k['ms_per_pass'], 1))"
done
cp /tmp/liborbx_keep.so $SO
