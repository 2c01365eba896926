#!/bin/bash
mkdir -p gpurun_out/r2m2
for r in 1 2 3; do
LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2m2/b_$r.json 2> gpurun_out/r2m2/b_$r.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2m2/b_$r.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("run $r: e2e %.0f init_ms %.1f fetch_ms %.1f" % (e["value"], e["init_ms"], e["fetch_fields_ms"]))
PY
grep "lbgpu trace" gpurun_out/r2m2/b_$r.err | awk '/type scan/{n++} n>=2' | head -n 22
done
