#!/bin/bash
# round 2, visit P2 (1 GPU): compute-sanitizer on the grid rebuild of the DEM neighbour table and on the DEM-coupled graph replay
mkdir -p gpurun_out/r2p2
cd tests
LBGPU_DEM_GRID=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python mini_run.py bed_dem 8 --dem > ../gpurun_out/r2p2/sanitizer_memcheck_bed_dem_grid.log 2>&1
echo "memcheck grid rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2p2/sanitizer_memcheck_bed_dem_grid.log | tr '\n' ' ')"
LBGPU_DEM_GRID=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python mini_run.py bed_dem 8 --dem > ../gpurun_out/r2p2/sanitizer_racecheck_bed_dem_grid.log 2>&1
echo "racecheck grid rc=$? : $(grep -E 'RACECHECK SUMMARY|^ok' ../gpurun_out/r2p2/sanitizer_racecheck_bed_dem_grid.log | tr '\n' ' ')"
LBGPU_DEM_GRID=1 timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python mini_run.py bed_dem 8 --dem > ../gpurun_out/r2p2/sanitizer_initcheck_bed_dem_grid.log 2>&1
echo "initcheck grid rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2p2/sanitizer_initcheck_bed_dem_grid.log | tr '\n' ' ')"
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python mini_run.py cfg4_mini 8 --run > ../gpurun_out/r2p2/sanitizer_initcheck_cfg4_mini.log 2>&1
echo "initcheck cfg4_mini rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2p2/sanitizer_initcheck_cfg4_mini.log | tr '\n' ' ')"
