#!/bin/bash
# round 2, visit C (1 GPU): whole GPU suite, L2-prefetch A/B, transfer trace
mkdir -p gpurun_out/r2c
( time timeout 1800 python -m pytest tests -q -m gpu ) > gpurun_out/r2c/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2c/pytest.log | tail -30
for pf in 0 370 740 1480 2960 5920; do
  LBGPU_PREFETCH=$pf timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2c/ab_pf$pf.json 2> gpurun_out/r2c/ab_pf$pf.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2c/ab_pf$pf.json").read().strip().splitlines()[-1])
print("prefetch=$pf", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel frac %.4f" % d["roofline"]["frac"], d["clocks"])
PY
done
for th in 4 16; do
LBGPU_COPY_THREADS=$th LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2c/bench20_t$th.json 2> gpurun_out/r2c/bench20_t$th.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c/bench20_t$th.json").read().strip().splitlines()[-1])
print("threads=$th bench20", "MLUPS %.0f" % d["value"], "frac %.4f" % d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["init_ms"], d["e2e"]["fetch_fields_ms"], d["clocks"])
PY
done
grep "lbgpu trace" gpurun_out/r2c/bench20_t16.err | tail -18
