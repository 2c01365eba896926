#!/bin/bash
# round 2, visit S (1 GPU): shim with the DEM on the device, whole GPU suite, smoke, default bench
mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_shim.py -x -q -m gpu > gpurun_out/r2s/pytest_shim.log 2>&1
echo "shim rc=$?"; tail -n 25 gpurun_out/r2s/pytest_shim.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s/pytest.log 2>&1
echo "all rc=$?"; tail -n 6 gpurun_out/r2s/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s/smoke.log 2>&1; tail -n 3 gpurun_out/r2s/smoke.log
