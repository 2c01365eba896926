#!/bin/bash
# round 2, visit W (1 GPU): pooled macro arrays / pinned stages, tabulated ghost masks: parity suite, then the default bench with the init trace
mkdir -p gpurun_out/r2w
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2w/pytest.log 2>&1
echo "all rc=$?"; tail -n 6 gpurun_out/r2w/pytest.log
LBGPU_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w/bench20.json 2> gpurun_out/r2w/bench20.err
grep "lbgpu trace" gpurun_out/r2w/bench20.err | grep -v Fetch | tail -n 14
python - <<PY
import json
d = json.loads(open("gpurun_out/r2w/bench20.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("bench20 MLUPS %.0f frac %.3f e2e %.0f init_ms %.1f fetch_ms %.1f" % (d["value"], d["roofline"]["frac"], e["value"], e["init_ms"], e["fetch_fields_ms"]))
for k, v in d.get("extra", {}).items():
    print("   ", k, "ms/step %.4f" % v["ms_per_step"], "frac %.3f whole %.3f" % (v["roofline"]["frac"], v["roofline"]["whole_step_frac"]))
PY
