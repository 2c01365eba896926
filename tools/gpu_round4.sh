#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q -k "drum or cfg4_mini or couette" > gpurun_out/pytest_drum.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_drum.log
tail -30 gpurun_out/pytest_drum.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=8 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -40 gpurun_out/pytest_full.log
