#!/bin/bash
# round 2, visit X2 (1 GPU): GPU suite on the final tree
mkdir -p gpurun_out/r2x2
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2x2/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 7 gpurun_out/r2x2/pytest.log
