#!/bin/bash
# round 2, visit F (1 GPU): launch lists of cfg4 / cfg5 / cfg3, ncu full of the cfg2 kernel with L2 prefetch
mkdir -p gpurun_out/r2f
export LBGPU_PREFETCH=740
for w in cfg4 cfg5 cfg3; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2f/launches_$w.csv python bench.py --workload $w --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2f/b_ncu_$w.log 2>&1
python tools/launch_table.py gpurun_out/r2f/launches_$w.csv | sort -t= -k2 | head -40
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 2 -f -o gpurun_out/r2f/prof_step_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2f/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/r2f/prof_step_cfg4 python bench.py --workload cfg4 --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2f/b_ncu4f.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 200 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2f/bench_cfg4.json 2>gpurun_out/r2f/bench_cfg4.err; tail -c 600 gpurun_out/r2f/bench_cfg4.json
ls -la gpurun_out/r2f
