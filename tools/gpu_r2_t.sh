#!/bin/bash
# round 2, visit T (1 GPU): what bounds the free-surface bulk launch (Bingham vs Newtonian on the same geometry); where lbGpuInit's time goes
mkdir -p gpurun_out/r2t
for w in cfg4 cfg4_newtonian; do
  timeout 600 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2t/bench_$w.json 2> gpurun_out/r2t/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k_step -s 30 -c 6 --csv --log-file gpurun_out/r2t/ncu_cfg4_newtonian.csv python bench.py --workload cfg4_newtonian --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2t/b_ncu.log 2>&1
LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2t/bench20_trace.json 2> gpurun_out/r2t/bench20_trace.err
grep "lbgpu trace" gpurun_out/r2t/bench20_trace.err | tail -n 40
python - <<PY
import json
for w in ("cfg4", "cfg4_newtonian"):
    d = json.loads(open("gpurun_out/r2t/bench_%s.json" % w).read().strip().splitlines()[-1])
    print(w, "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "frac %.3f" % d["roofline"]["frac"], "whole %.3f" % d["roofline"]["whole_step_frac"])
PY
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2t/ncu_cfg4_newtonian.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
