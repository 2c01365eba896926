#!/bin/bash
# round 2, visit A (1 GPU): the GPU test suite incl. the 1000-step / full-size fixtures, bench with the transfer trace
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a/gpu.txt
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2a/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r2a/bench20.json 2> gpurun_out/r2a/bench20.err
tail -c 1500 gpurun_out/r2a/bench20.json
grep "lbgpu trace" gpurun_out/r2a/bench20.err | tail -40
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench200_extra.json 2> gpurun_out/r2a/bench200_extra.err
tail -c 3000 gpurun_out/r2a/bench200_extra.json
