#!/bin/bash
python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "exit $?"; tail -3 gpurun_out/bench_short.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_short.json'))
print({k:round(d[k],3) for k in ("value","sq_frames_per_s","gr_frames_per_s","sq_default_grid_frames_per_s") if k in d}, round(d['roofline']['frac'],3), round(d['e2e']['value'],3))
PY
