#!/bin/bash
# usage: gpu_prof.sh <kernel-regex> <outfile-stem> [bench args...]   -- tests, plain bench, then one ncu --set full capture
K=$1; O=$2; shift 2
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench_plain.json 2> gpurun_out/plain.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_plain.json'))
print({k:d[k] for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','gr_pairs_per_s')}, d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value'])
PY
python bench.py --steps 1 --warmup 1 --frames 1 --no-cpu "$@" > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/$O python bench.py --steps 1 --warmup 1 --frames 1 --no-cpu "$@" > gpurun_out/ncu_$O.log 2>&1
echo "ncu exit $?"
