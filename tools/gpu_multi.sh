#!/bin/bash
# multi-GPU visit: N = number of visible GPUs
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slabs.py -m gpu -q -k "two_processes" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python bench.py --gpus $N --steps 500 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 900 python bench.py --gpus $N --workload cfg5 --steps 100 --warmup 5 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; echo "bench cfg5 N=$N rc=$?"; cat gpurun_out/bench_cfg5_n$N.json; tail -5 gpurun_out/bench_cfg5_n$N.err
