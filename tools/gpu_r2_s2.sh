#!/bin/bash
mkdir -p gpurun_out/r2s2
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s -k "periodic" > gpurun_out/r2s2/pytest_dem.log 2>&1
echo "rc=$?"; tail -n 14 gpurun_out/r2s2/pytest_dem.log
