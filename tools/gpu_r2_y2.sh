#!/bin/bash
# round 2, visit Y2 (1 GPU): the rest of the GPU suite after the shim rebuild (the suite stopped at the first failure)
mkdir -p gpurun_out/r2y2
( time timeout 1200 python -m pytest tests/test_gpu_shim.py tests/test_gpu_slabs.py -q -m gpu ) > gpurun_out/r2y2/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2y2/pytest.log
