#!/bin/bash
# round 2, visit Q2 (1 GPU): clusters on the device: DEM tests
mkdir -p gpurun_out/r2q2
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s > gpurun_out/r2q2/pytest_dem.log 2>&1
echo "dem rc=$?"; tail -n 25 gpurun_out/r2q2/pytest_dem.log
