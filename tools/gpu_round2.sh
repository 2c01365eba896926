#!/bin/bash
# GPU visit: parity tests, bench on cfg2/3/4, launch list of a free-surface step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
for w in cfg3 cfg4; do
timeout 600 python bench.py --workload $w --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu4.log 2>&1
ls -la gpurun_out
