#!/bin/bash
# round 2, visit A2 (1 GPU): device-memory cache: parity suite (dozens of engines per process reuse each other's memory), e2e repeated
mkdir -p gpurun_out/r2a2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a2/pytest.log 2>&1
echo "all rc=$?"; tail -n 6 gpurun_out/r2a2/pytest.log
for r in 1 2 3 4; do
  LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2a2/b_$r.json 2> gpurun_out/r2a2/b_$r.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2a2/b_$r.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("run $r: value %.0f e2e %.0f init_ms %.1f fetch_ms %.1f" % (d["value"], e["value"], e["init_ms"], e["fetch_fields_ms"]))
PY
done
LBGPU_CACHE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2a2/b_nocache.json 2> gpurun_out/r2a2/b_nocache.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2a2/b_nocache.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("no cache: value %.0f e2e %.0f init_ms %.1f fetch_ms %.1f" % (d["value"], e["value"], e["init_ms"], e["fetch_fields_ms"]))
PY
