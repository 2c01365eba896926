#!/bin/bash
# final visit of the round: bench (both arms) on cfg2, bench on cfg3-5, launch lists, ncu full of the cfg2 and cfg5 step kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python bench.py --steps 1000 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
for w in cfg3 cfg4 cfg5; do
timeout 900 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
cat gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_cfg2.log 2>&1
for w in cfg3 cfg4 cfg5; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 600 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 2 -f -o gpurun_out/prof_step python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/prof_step_cfg5 python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu5f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/prof_step_cfg4 python bench.py --workload cfg4 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu4f.log 2>&1
ls -la gpurun_out | head -50
