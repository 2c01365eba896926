#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for c in cfg3 cfg4 cfg5; do
bash tools/ab.sh $c 100 hybird_b200/_ab/lib_c4s4.so hybird_b200/_ab/lib_c5s5.so hybird_b200/_ab/lib_c4s5.so hybird_b200/_ab/lib_c5s4.so > gpurun_out/ab_$c.log 2>&1
cat gpurun_out/ab_$c.log
done
