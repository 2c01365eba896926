#!/bin/bash
mkdir -p gpurun_out/r2v2
timeout 600 python tools/cfg5_dem_probe.py 130 > gpurun_out/r2v2/probe.log 2>&1; tail -n 24 gpurun_out/r2v2/probe.log
