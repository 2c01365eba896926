#!/bin/bash
# round 2, visit L (1 GPU): list builds without a scan launch + CUDA graph of the free-surface cycle: parity, then A/B
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2l/pytest.log 2>&1
echo "all rc=$?"; tail -n 8 gpurun_out/r2l/pytest.log
for w in cfg1 cfg4 cfg5 cfg3; do
  for gr in 1 0; do
    LBGPU_GRAPH=$gr timeout 600 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2l/bench_${w}_g$gr.json 2> gpurun_out/r2l/bench_${w}_g$gr.err
  done
done
python - <<PY
import json
for w in ("cfg1", "cfg4", "cfg5", "cfg3"):
    for g in (1, 0):
        f = "gpurun_out/r2l/bench_%s_g%d.json" % (w, g)
        try:
            d = json.loads(open(f).read().strip().splitlines()[-1])
            print(w, "graph", g, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"], "launches", d["gpu_launches"])
        except Exception as e:
            print(w, g, "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
