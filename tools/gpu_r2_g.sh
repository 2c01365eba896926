#!/bin/bash
# round 2, visit G (1 GPU): GPU suite + default bench with the extra block (hoisted scalar loads, gas-only mixed tiles)
mkdir -p gpurun_out/r2g
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r2g/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2g/pytest.log | tail -30
timeout 900 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2g/bench.json").read().strip().splitlines()[-1])
print("cfg2", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel frac %.4f" % d["roofline"]["frac"], "e2e %.0f init %.1f fetch %.1f" % (d["e2e"]["value"], d["e2e"]["init_ms"], d["e2e"]["fetch_fields_ms"]), d["clocks"])
for k, v in d.get("extra", {}).items():
    if "error" in v: print(k, v); continue
    print("   ", k, "MLUPS %.0f" % v["value"], "ms/step %.4f" % v["ms_per_step"], "kernel frac %.3f whole-step frac %.3f" % (v["roofline"]["frac"], v["roofline"]["whole_step_frac"]), "launches/step", v["launches_per_step"], "init_s", v["init_s"])
PY
