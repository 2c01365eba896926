#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_init.py tests/test_gpu_shim.py -m gpu -q > gpurun_out/pytest_dinit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dinit.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_dinit.log | cut -c1-220 | head -40
