#!/bin/bash
mkdir -p gpurun_out/r2u2
timeout 600 python tools/cfg5_dem_probe.py 10 > gpurun_out/r2u2/probe.log 2>&1; tail -n 16 gpurun_out/r2u2/probe.log
echo ---- sync flood fill
LBGPU_FLOOD_GENS=0 timeout 600 python tools/cfg5_dem_probe.py 6 > gpurun_out/r2u2/probe_sync.log 2>&1; tail -n 9 gpurun_out/r2u2/probe_sync.log
