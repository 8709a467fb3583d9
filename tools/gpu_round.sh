#!/bin/bash
# One GPU-box visit: parity tests, diagnostics, bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/diag_sq.py > gpurun_out/diag_sq.log 2>&1; tail -40 gpurun_out/diag_sq.log
python bench.py --steps 3 --warmup 3 --no-cpu --sq-algo direct > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench direct exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_direct.json','gpurun_out/bench.json'):
    try:
        d=json.load(open(f))
        print({k:d[k] for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','gr_pairs_per_s')}, d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value'])
    except Exception as e: print(f, e)
PY
tail -3 gpurun_out/bench.err
