#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/nufft_launches.csv python tools/nufft_check.py 100000 161 20 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nufft_spread|nufft_fft" -c 4 -f -o gpurun_out/nufft_full python tools/nufft_check.py 100000 161 20 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
