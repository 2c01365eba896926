#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
