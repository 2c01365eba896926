#!/bin/bash
mkdir -p gpurun_out
cd tests && timeout 300 python golden_divergence.py drum_bingham 200 > ../gpurun_out/div_drum_bingham.log 2>&1; timeout 300 python golden_divergence.py drum_mini 300 > ../gpurun_out/div_drum_mini.log 2>&1; cd ..
tail -12 gpurun_out/div_drum_bingham.log | cut -c1-400; tail -5 gpurun_out/div_drum_mini.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for w in cfg3 cfg4 cfg5; do
timeout 900 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --workload cfg3 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/prof_step_cfg5 python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu5f.log 2>&1
