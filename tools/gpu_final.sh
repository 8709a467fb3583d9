#!/bin/bash
# round-end evidence: parity tests, smoke, full bench (with the CPU baseline leg), reference arm, launch list, ncu captures
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
python tools/extra_bench.py > gpurun_out/extra.log 2>&1; echo "extra exit $?"
(python tools/rdf_time.py 8 4; python tools/sq_time.py 8 4; python tools/rdf_tri_time.py; MDC_RDF_TRI_FILTER=0 python tools/rdf_tri_time.py; MDC_RDF_REL=0 python tools/rdf_time.py 8 4; MDC_RDF_L32=0 python tools/rdf_time.py 8 4) > gpurun_out/alone.log 2>&1; cat gpurun_out/alone.log
[ -x build/atoms_probe ] && ./build/atoms_probe > gpurun_out/atoms_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --frames 8 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rdf_pairs_fast" -s 3 -c 1 -f -o gpurun_out/rdf_full python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-extras --one-stream > gpurun_out/ncu_rdf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nufft_records|nufft_spread|nufft_fft" -s 5 -c 5 -f -o gpurun_out/nufft_full python tools/nufft_check.py 100000 161 20 > gpurun_out/ncu_nufft.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','sq_frames_per_s','gr_frames_per_s','sq_default_grid_frames_per_s','gpu_launches','parity_check')}, d['e2e']['value'])
for k in ('roofline','roofline_sq','roofline_sfu_direct','roofline_fp32_factored'):
    r=d.get(k)
    if r: print(k, r['kernel'][:40], r['bound'], round(r['achieved'],1), round(r['peak'],1), r['unit'], round(r['frac'],3))
print(d.get('cpu_baseline'))
PY
