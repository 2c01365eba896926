#!/bin/bash
mkdir -p gpurun_out
for g in 0 4 8 16; do
echo "== LBGPU_FS_GRID=$g"
LBGPU_FS_GRID=$g python tests/quick_bench.py cfg4 150 2>&1 | grep -E "ms/step" | tail -2
LBGPU_FS_GRID=$g python tests/quick_bench.py cfg5 40 2>&1 | grep -E "ms/step" | tail -2
done > gpurun_out/ab_fsgrid.log 2>&1
cat gpurun_out/ab_fsgrid.log
