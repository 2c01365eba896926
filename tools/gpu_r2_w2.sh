#!/bin/bash
# round 2, visit W2 (1 GPU): the driver's round-end sequence on the final tree (GPU suite, smoke, both bench arms) + cfg5 with the DEM in the loop
mkdir -p gpurun_out/r2w2
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2w2/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 7 gpurun_out/r2w2/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w2/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2w2/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w2/bench20.json 2> gpurun_out/r2w2/bench20.err; echo "bench rc=$?"
timeout 600 python bench.py --workload cfg5_dem --steps 100 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2w2/bench_cfg5_dem.json 2> gpurun_out/r2w2/bench_cfg5_dem.err; echo "cfg5_dem rc=$?"
python - <<PY
import json
for f in ("bench20", "bench_cfg5_dem"):
    try:
        d = json.loads(open("gpurun_out/r2w2/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["config"].get("particles"), "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], "step_value %.0f" % d["e2e"]["step_value"], "launches", d["gpu_launches"])
        for k, v in d.get("extra", {}).items():
            print("   ", k, v.get("dem", v.get("error")), ("ms/step %.4f frac %.3f whole %.3f" % (v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["whole_step_frac"])) if "ms_per_step" in v else "", v.get("phase_ms", {}).get("dem"))
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/r2w2/%s.err" % f).read()[-600:])
PY
