#!/bin/bash
# round 2, visit N2 (1 GPU): grid rebuild of the DEM neighbour table vs all pairs; DEM timing with the grid
mkdir -p gpurun_out/r2n2
timeout 900 python -m pytest tests/test_gpu_dem.py -x -q -m gpu > gpurun_out/r2n2/pytest_dem.log 2>&1
echo "dem rc=$?"; tail -n 6 gpurun_out/r2n2/pytest_dem.log
for gr in 0 1; do
LBGPU_DEM_GRID=$gr timeout 600 python tools/dem_bench.py 20000 300 > gpurun_out/r2n2/dem_bench_20000_g$gr.json 2> gpurun_out/r2n2/dem_bench_20000_g$gr.err; cat gpurun_out/r2n2/dem_bench_20000_g$gr.json
done
LBGPU_DEM_GRID=1 timeout 600 python tools/dem_bench.py 200000 100 > gpurun_out/r2n2/dem_bench_200000_g1.json 2> gpurun_out/r2n2/dem_bench_200000_g1.err; cat gpurun_out/r2n2/dem_bench_200000_g1.json; tail -n 2 gpurun_out/r2n2/dem_bench_200000_g1.err
