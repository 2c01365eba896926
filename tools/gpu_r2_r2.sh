#!/bin/bash
# round 2, visit R2 (1 GPU): cfg5 with the DEM in the loop on the device
mkdir -p gpurun_out/r2r2
timeout 600 python bench.py --workload cfg5_dem --steps 100 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2r2/bench_cfg5_dem.json 2> gpurun_out/r2r2/bench_cfg5_dem.err
echo "rc=$?"; tail -n 5 gpurun_out/r2r2/bench_cfg5_dem.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2r2/bench_cfg5_dem.json").read().strip().splitlines()[-1])
    print(d["config"].get("particles"), "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e)
PY
