#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
