#!/bin/bash
# A/B of an environment switch on one box: GPU tests, then the bench with and without it.  usage: gpu_ab.sh VAR=VALUE
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for tag in on off; do
  if [ $tag = off ]; then export "$1"; fi
  python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench $tag exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('$tag', {k:round(d[k],2) for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s')}, d['e2e']['value'], d.get('parity_check'))
PY
  tail -2 gpurun_out/bench_$tag.err
done
