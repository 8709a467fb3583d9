#!/bin/bash
# usage: gpu_bench.sh [bench args...]   -- one bench run at N = 1, line to gpurun_out/bench.json, a digest to stdout
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
except Exception as e:
    print("no line:", e); raise SystemExit
keys=('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','sq_default_grid_frames_per_s','gpu_launches','parity_check')
print({k:d.get(k) for k in keys})
print('e2e', d['e2e']['value'], d['e2e']['seconds'], 'capi', d['e2e_capi']['value'])
print('timing', d['timing'])
print({k:v for k,v in d.items() if k.startswith(('sq_max','gr_slab','gr_in','classes','allreduce','parity'))})
for k in ('roofline','roofline_sq','roofline_sfu_direct','roofline_fp32_factored'):
    r=d.get(k)
    if r: print(k, r['kernel'][:40], r['bound'], round(r['achieved'],1), round(r['peak'],1), r['unit'], round(r['frac'],3), r.get('traffic'), r.get('issue_active'), r.get('slots_per_pair'))
print(d.get('cpu_baseline'))
print(d.get('clocks'))
PY
