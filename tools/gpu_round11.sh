#!/bin/bash
mkdir -p gpurun_out
for pf in 0 5 4 3 5 0; do
echo "== LBGPU_PF=$pf"
LBGPU_PF=$pf python tests/quick_bench.py cfg2 500 2>&1 | grep -E "ms/step|rror" | tail -2
done > gpurun_out/ab_pf.log 2>&1
cat gpurun_out/ab_pf.log
LBGPU_PF=5 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py -m gpu -q -k "cfg2 or channel_oblique or box_noforce or periodic_all or launches" > gpurun_out/pytest_pf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pf.log
tail -6 gpurun_out/pytest_pf.log
python bench.py --steps 300 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-300
