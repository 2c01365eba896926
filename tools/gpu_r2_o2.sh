#!/bin/bash
mkdir -p gpurun_out/r2o2
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu > gpurun_out/r2o2/pytest_dem.log 2>&1
echo "dem rc=$?"; tail -n 4 gpurun_out/r2o2/pytest_dem.log
