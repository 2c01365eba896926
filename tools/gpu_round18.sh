#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_device_init.py -m gpu -q > gpurun_out/pytest_dinit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dinit.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_dinit.log | cut -c1-220 | head -20
for w in cfg2 cfg4 cfg5; do
timeout 900 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_di_$w.json 2> gpurun_out/bench_di_$w.err; echo "bench $w rc=$?"
cut -c1-200 gpurun_out/bench_di_$w.json; tail -3 gpurun_out/bench_di_$w.err
done
