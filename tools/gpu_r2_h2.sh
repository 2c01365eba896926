#!/bin/bash
# round 2, visit H2 (1 GPU): the driver's round-end sequence on the final tree: GPU suite, smoke, reference arm, default bench
mkdir -p gpurun_out/r2h2
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2h2/pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/r2h2/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h2/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2h2/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r2h2/bench_ref.json 2> gpurun_out/r2h2/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2h2/bench.json 2> gpurun_out/r2h2/bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("bench_ref", "bench"):
    try:
        d = json.loads(open("gpurun_out/r2h2/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "steps", d["steps"], "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac", (d.get("roofline") or {}).get("frac"), "e2e %.0f" % d["e2e"]["value"], d.get("clocks"), "launches", d.get("gpu_launches"))
        for k, v in d.get("extra", {}).items():
            print("   ", k, v.get("dem"), "ms/step %.4f" % v["ms_per_step"], "frac %.3f whole %.3f" % (v["roofline"]["frac"], v["roofline"]["whole_step_frac"]))
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/r2h2/%s.err" % f).read()[-800:])
PY
