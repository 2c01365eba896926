#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for c in cfg5 cfg1; do
for wp in 0 1; do
echo "== $c LBGPU_WALL_PUSH=$wp"
LBGPU_WALL_PUSH=$wp python tests/quick_bench.py $c 100 2>&1 | grep -E "ms/step|rror" | tail -2
done; done > gpurun_out/ab_wallpush2.log 2>&1
cat gpurun_out/ab_wallpush2.log
