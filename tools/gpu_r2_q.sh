#!/bin/bash
# round 2, visit Q (1 GPU): A/B of the block-wide L2 prefetch for the tile-list launches (cfg4, cfg5)
mkdir -p gpurun_out/r2q
for w in cfg4 cfg5; do
  for pf in 0 370 740 1480; do
    LBGPU_PREFETCH_TILES=$pf timeout 600 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2q/bench_${w}_pf$pf.json 2> gpurun_out/r2q/bench_${w}_pf$pf.err
  done
done
python - <<PY
import json
for w in ("cfg4", "cfg5"):
    for pf in (0, 370, 740, 1480):
        f = "gpurun_out/r2q/bench_%s_pf%d" % (w, pf)
        try:
            d = json.loads(open(f + ".json").read().strip().splitlines()[-1])
            print(w, "pf", pf, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"])
        except Exception as e:
            print(w, pf, "failed", e, open(f + ".err").read()[-600:])
PY
