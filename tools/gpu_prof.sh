#!/bin/bash
# tests + ncu capture of the S(q) kernels on a short bench run
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 1 --warmup 1 --frames 1 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sq_lattice_factored -s 1 -c 1 -f -o gpurun_out/prof_sq_factored python bench.py --steps 1 --warmup 1 --frames 1 --no-cpu > gpurun_out/ncu1.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu1.log
