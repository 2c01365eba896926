#!/bin/bash
# round 2, visit F2 (1 GPU): DEM-coupled cycle in the graph, cfg3 at full size with the DEM on the device, bench cfg3 (runDem)
mkdir -p gpurun_out/r2f2
timeout 1500 python -m pytest tests/test_gpu_dem.py -x -q -m gpu -s > gpurun_out/r2f2/pytest_dem.log 2>&1
echo "dem rc=$?"; tail -n 12 gpurun_out/r2f2/pytest_dem.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_dem.py > gpurun_out/r2f2/pytest.log 2>&1
echo "rest rc=$?"; tail -n 4 gpurun_out/r2f2/pytest.log
for gr in 1 0; do
LBGPU_GRAPH=$gr timeout 600 python bench.py --workload cfg3 --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2f2/bench_cfg3_g$gr.json 2> gpurun_out/r2f2/bench_cfg3_g$gr.err
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f2/bench20.json 2> gpurun_out/r2f2/bench20.err
python - <<PY
import json
for f in ("bench_cfg3_g1", "bench_cfg3_g0", "bench20"):
    try:
        d = json.loads(open("gpurun_out/r2f2/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["config"].get("particles"), "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "whole %.3f" % d["roofline"]["whole_step_frac"], "launches", d["gpu_launches"], "e2e %.0f" % d["e2e"]["value"], "init_ms %.0f" % d["e2e"]["init_ms"])
        for k, v in d.get("extra", {}).items():
            print("   ", k, v.get("dem"), "ms/step %.4f" % v["ms_per_step"], "launches/step", v["launches_per_step"], v.get("phase_ms"))
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/r2f2/%s.err" % f).read()[-800:])
PY
