#!/bin/bash
# A/B timing of alternative builds of liblbgpu.so on one workload:  tools_ab.sh CASE STEPS lib1 lib2 ...
case=$1; steps=$2; shift 2
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" = "tiles" ]; then LBGPU_TILES=1 python tests/quick_bench.py $case $steps 2>&1 | grep -E "MLUPS|rror"
  elif [ "$lib" = "default" ]; then python tests/quick_bench.py $case $steps 2>&1 | grep -E "MLUPS|rror"
  else LBGPU_LIB=$lib python tests/quick_bench.py $case $steps 2>&1 | grep -E "MLUPS|rror"; fi
done
