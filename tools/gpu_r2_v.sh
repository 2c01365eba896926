#!/bin/bash
# round 2, visit V (1 GPU): host-side copy threads of the staged transfers (lbGpuInit / lbGpuFetchFields from pageable arrays)
mkdir -p gpurun_out/r2v
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)" 
for t in 4 8 12 16; do
  LBGPU_COPY_THREADS=$t LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2v/bench_t$t.json 2> gpurun_out/r2v/bench_t$t.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2v/bench_t$t.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("threads $t e2e %.0f init_ms %.1f fetch_ms %.1f" % (e["value"], e["init_ms"], e["fetch_fields_ms"]))
PY
  grep "lbgpu trace" gpurun_out/r2v/bench_t$t.err | grep -E "up|Fetch" | tail -n 8
done
