#!/bin/bash
mkdir -p gpurun_out/r2t2
LBGPU_DEM_GRID=1 timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s -k "periodic or coupled_cycle_on_device" > gpurun_out/r2t2/pytest_dem_grid.log 2>&1
echo "grid rc=$?"; tail -n 14 gpurun_out/r2t2/pytest_dem_grid.log
