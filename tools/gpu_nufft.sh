#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "nufft or patch_edges" > gpurun_out/pytest_nufft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nufft.log
tail -5 gpurun_out/pytest_nufft.log
timeout 600 python tools/nufft_check.py 100000 161 20 > gpurun_out/nufft_check_161.log 2>&1; tail -12 gpurun_out/nufft_check_161.log
timeout 600 python tools/nufft_check.py 100000 32 0 > gpurun_out/nufft_check_32.log 2>&1; tail -12 gpurun_out/nufft_check_32.log
timeout 600 python tools/nufft_check.py 200000 201 20 > gpurun_out/nufft_check_201.log 2>&1; tail -12 gpurun_out/nufft_check_201.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/nufft_launches.csv python tools/nufft_check.py 100000 161 20 > gpurun_out/ncu1.log 2>&1
grep -E "nufft|pair_accum" gpurun_out/nufft_launches.csv | tail -9 | cut -d, -f5,15- 
