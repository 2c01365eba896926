#!/bin/bash
# round 2, visit L2 (1 GPU): launch lists of cfg3-cfg5 and the bench lines on the final tree
mkdir -p gpurun_out/r2l2
for w in cfg3 cfg4 cfg5; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2l2/launches_$w.csv python bench.py --workload $w --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2l2/b_ncu_$w.log 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l2/bench20.json 2> gpurun_out/r2l2/bench20.err
for w in cfg1 cfg3 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2l2/bench_$w.json 2> gpurun_out/r2l2/bench_$w.err; done
python - <<PY
import json
for f in ("bench20", "bench_cfg1", "bench_cfg3", "bench_cfg4", "bench_cfg5"):
    try:
        d = json.loads(open("gpurun_out/r2l2/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"], "e2e %.0f" % d["e2e"]["value"], "init_ms %.0f" % d["e2e"]["init_ms"], d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
