#!/bin/bash
# round 2, visit K (1 GPU): device-side DEM tests, then the whole GPU suite
mkdir -p gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s > gpurun_out/r2k/pytest_dem.log 2>&1
echo "dem rc=$?"; tail -n 25 gpurun_out/r2k/pytest_dem.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k/pytest.log 2>&1
echo "all rc=$?"; tail -n 5 gpurun_out/r2k/pytest.log
