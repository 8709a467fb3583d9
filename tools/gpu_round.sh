#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','sq_default_grid_frames_per_s','gpu_launches')}, d['e2e'])
for k in ('roofline','roofline_sq','roofline_sfu_direct','roofline_fp32_factored'):
    r=d[k]; print(k, r['kernel'][:40], r['bound'], round(r['achieved'],1), round(r['peak'],1), r['unit'], round(r['frac'],3))
print(d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench.err
timeout 300 python tools/nufft_check.py 100000 161 20 > gpurun_out/nufft_check_161.log 2>&1; grep -E "FP64 reference|nufft " gpurun_out/nufft_check_161.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nufft_spread|nufft_fft" -c 4 -f -o gpurun_out/nufft_full python tools/nufft_check.py 100000 161 20 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
