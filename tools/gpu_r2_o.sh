#!/bin/bash
# round 2, visit O (1 GPU): deferred flood fill + graph of coupled cycles: parity suite, A/B on cfg3 / cfg5, DEM sub-step timing, default bench
mkdir -p gpurun_out/r2o
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2o/pytest.log 2>&1
echo "all rc=$?"; tail -n 8 gpurun_out/r2o/pytest.log
for w in cfg3 cfg5; do
  for gr in 1 0; do
    LBGPU_GRAPH=$gr timeout 600 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2o/bench_${w}_g$gr.json 2> gpurun_out/r2o/bench_${w}_g$gr.err
  done
done
LBGPU_GRAPH=0 LBGPU_FLOOD_GENS=0 timeout 600 python bench.py --workload cfg3 --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2o/bench_cfg3_sync.json 2> gpurun_out/r2o/bench_cfg3_sync.err
timeout 600 python tools/dem_bench.py 20000 300 > gpurun_out/r2o/dem_bench_20000.json 2> gpurun_out/r2o/dem_bench_20000.err
timeout 600 python tools/dem_bench.py 2000 300 > gpurun_out/r2o/dem_bench_2000.json 2> gpurun_out/r2o/dem_bench_2000.err
cat gpurun_out/r2o/dem_bench_*.json; tail -n 3 gpurun_out/r2o/dem_bench_20000.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o/bench20.json 2> gpurun_out/r2o/bench20.err
python - <<PY
import json
for f in ("bench_cfg3_g1", "bench_cfg3_g0", "bench_cfg3_sync", "bench_cfg5_g1", "bench_cfg5_g0", "bench20"):
    try:
        d = json.loads(open("gpurun_out/r2o/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"], "launches", d["gpu_launches"], "e2e %.0f" % d["e2e"]["value"])
        for k, v in d.get("extra", {}).items():
            print("   ", k, {kk: v[kk] for kk in v if kk not in ("workload", "roofline", "lattice")}, "frac %.3f whole %.3f" % (v["roofline"]["frac"], v["roofline"]["whole_step_frac"]) if "roofline" in v else "")
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/r2o/%s.err" % f).read()[-800:])
PY
