#!/bin/bash
# round 2, visit X (1 GPU): end-to-end job with resident host arrays, per-array upload trace
mkdir -p gpurun_out/r2x
LBGPU_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x/bench20.json 2> gpurun_out/r2x/bench20.err
grep "lbgpu trace" gpurun_out/r2x/bench20.err | tail -n 24
python - <<PY
import json
d = json.loads(open("gpurun_out/r2x/bench20.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("bench20 MLUPS %.0f frac %.3f e2e %.0f init_ms %.1f fetch_ms %.1f" % (d["value"], d["roofline"]["frac"], e["value"], e["init_ms"], e["fetch_fields_ms"]))
PY
