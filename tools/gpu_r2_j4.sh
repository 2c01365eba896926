#!/bin/bash
# round 2, visit J4 (4 GPUs): four-rank parity tests
mkdir -p gpurun_out/r2j4
( time timeout 1200 python -m pytest tests -q -m gpu -k "four_processes" ) > gpurun_out/r2j4/pytest_4gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j4/pytest_4gpu.log
tail -n 12 gpurun_out/r2j4/pytest_4gpu.log
