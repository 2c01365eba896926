#!/bin/bash
# round 2, visit U (1 GPU): windows anchored at the runs of fluid cells instead of aligned tiles: parity suite, then A/B
mkdir -p gpurun_out/r2u
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2u/pytest.log 2>&1
echo "all rc=$?"; tail -n 8 gpurun_out/r2u/pytest.log
for w in cfg4 cfg5 cfg1; do
  for wd in 1 0; do
    LBGPU_WINDOWS=$wd timeout 600 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2u/bench_${w}_w$wd.json 2> gpurun_out/r2u/bench_${w}_w$wd.err
  done
done
python - <<PY
import json
for w in ("cfg4", "cfg5", "cfg1"):
    for g in (1, 0):
        f = "gpurun_out/r2u/bench_%s_w%d" % (w, g)
        try:
            d = json.loads(open(f + ".json").read().strip().splitlines()[-1])
            print(w, "windows", g, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"])
        except Exception as e:
            print(w, g, "failed", e, open(f + ".err").read()[-600:])
PY
