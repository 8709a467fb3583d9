#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "nufft or patch_edges" > gpurun_out/pytest_nufft.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_nufft.log
tail -25 gpurun_out/pytest_nufft.log
timeout 600 python tools/nufft_check.py 100000 161 20 > gpurun_out/nufft_check_161.log 2>&1; tail -20 gpurun_out/nufft_check_161.log
MDC_NUFFT_SORT=0 timeout 600 python tools/nufft_check.py 100000 161 20 2>&1 | head -3 > gpurun_out/nufft_check_161_nosort.log; cat gpurun_out/nufft_check_161_nosort.log
timeout 600 python tools/nufft_check.py 100000 32 0 > gpurun_out/nufft_check_32.log 2>&1; tail -12 gpurun_out/nufft_check_32.log
