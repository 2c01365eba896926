#!/bin/bash
mkdir -p gpurun_out
cd tests && timeout 300 python golden_divergence.py drum_bingham 200 > ../gpurun_out/div_drum_bingham.log 2>&1; timeout 300 python golden_divergence.py drum_mini 300 > ../gpurun_out/div_drum_mini.log 2>&1; cd ..
head -3 gpurun_out/div_drum_bingham.log | cut -c1-400; head -3 gpurun_out/div_drum_mini.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
for w in cfg3 cfg4 cfg5; do
timeout 900 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
for w in cfg3 cfg4 cfg5; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 600 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$w.log 2>&1
done
