#!/bin/bash
mkdir -p gpurun_out/r2z2
timeout 900 python -m pytest tests/test_gpu_dem.py tests/test_gpu_checkpoint.py -x -q -m gpu > gpurun_out/r2z2/pytest_dem.log 2>&1
echo "rc=$?"; tail -n 5 gpurun_out/r2z2/pytest_dem.log
