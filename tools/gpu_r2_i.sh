#!/bin/bash
# round 2, visit I (1 GPU): compute-sanitizer memcheck + racecheck on three minis, the list-growth test, launch list + ncu full of cfg2
mkdir -p gpurun_out/r2i
cd tests
for c in cfg5_mini couette_dyn drum_mini; do
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python mini_run.py $c 6 > ../gpurun_out/r2i/sanitizer_${tool}_$c.log 2>&1
    echo "$tool $c rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|^ok' ../gpurun_out/r2i/sanitizer_${tool}_$c.log | tr '\n' ' ')"
  done
done
cd ..
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "grow_ahead or sphere_fixed or cfg4_mini" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2i/b_ncu_cfg2.log 2>&1
python tools/launch_table.py gpurun_out/r2i/launches_cfg2.csv | tail -8
