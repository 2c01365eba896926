#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=8 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -25 gpurun_out/pytest_full.log
for w in cfg3 cfg5; do
timeout 900 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1
