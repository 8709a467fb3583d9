#!/bin/bash
# parity tests + a short bench (no CPU baseline leg)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','sq_default_grid_frames_per_s','gpu_launches')}, d['e2e']['value'])
for k in ('roofline','roofline_sq'):
    r=d[k]; print(k, r['kernel'][:40], r['bound'], round(r['achieved'],1), round(r['peak'],1), r['unit'], round(r['frac'],3))
PY
tail -3 gpurun_out/bench.err
