#!/bin/bash
# pyramid A/B: parity of the batch path, then the 512-frame bench with the tiled and the streaming kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "pyramid_streaming" 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do
  ORBX_PYR_PIPE=$v python bench.py --steps 10 --warmup 3 --no-cfg4 --no-knn2 --no-other --no-cpu 2> gpurun_out/pyr_$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('pipe=$v', d['value'], d['ms_per_step'], {k: round(v['ms_per_step'], 3) if isinstance(v, dict) and 'ms_per_step' in v else v for k, v in d.get('stages', {}).items()})"
done
for n in 6 12 16; do
  ORBX_PYR_PIPE_CTAS=$n python bench.py --steps 10 --warmup 3 --no-cfg4 --no-knn2 --no-other --no-cpu 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('ctas/sm=$n', d['value'], d['ms_per_step'], {k: round(v['ms_per_step'], 3) if isinstance(v, dict) and 'ms_per_step' in v else v for k, v in d.get('stages', {}).items()})"
done
