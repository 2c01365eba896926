#!/bin/bash
mkdir -p gpurun_out/r2g2
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s -k "cfg3_at_full_size" > gpurun_out/r2g2/pytest_dem.log 2>&1
echo "rc=$?"; tail -n 8 gpurun_out/r2g2/pytest_dem.log
