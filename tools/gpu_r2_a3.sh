#!/bin/bash
# round 2, visit A3 (2 GPUs): the DEM on two ranks, with and without periodic boundaries
mkdir -p gpurun_out/r2a3
( time timeout 600 python -m pytest tests/test_gpu_dem.py -q -m gpu -k "two_gpus" ) > gpurun_out/r2a3/pytest_2gpu_dem.log 2>&1
echo "rc=$?"; tail -n 8 gpurun_out/r2a3/pytest_2gpu_dem.log
