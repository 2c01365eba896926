#!/bin/bash
# round 2, visit Z (1 GPU): why lbGpuInit from host arrays takes 60 ms on one box and 400+ on another: stage size / pinned flags, repeated
mkdir -p gpurun_out/r2z
free -g | head -n 2; nproc
run() {
  tag=$1; shift
  for r in 1 2; do
    env "$@" LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2z/b_${tag}_$r.json 2> gpurun_out/r2z/b_${tag}_$r.err
    python - <<PY
import json, re
d = json.loads(open("gpurun_out/r2z/b_${tag}_$r.json").read().strip().splitlines()[-1])
e = d["e2e"]
tr = [l for l in open("gpurun_out/r2z/b_${tag}_$r.err") if "lbgpu trace" in l]
# the e2e handle is the second lbGpuInit of the process
idx = [i for i, l in enumerate(tr) if "type scan" in l]
seg = tr[idx[1]:] if len(idx) > 1 else tr
keys = ("population buffers allocated", "macroscopic arrays allocated", "allocations + memsets", "ghost lists: built", "type up", "solidIndex up", "n up", "mass up", "visc up", "u up", "ghosts, static")
vals = []
for k in keys:
    m = [re.search(r"([0-9.]+) ms \(total", l).group(1) for l in seg if k in l]
    vals.append(m[0] if m else "-")
print("${tag} run $r: e2e %.0f init_ms %.1f fetch_ms %.1f | " % (e["value"], e["init_ms"], e["fetch_fields_ms"]) + " ".join(vals))
PY
  done
}
run default LBGPU_X=0
run stage16 LBGPU_STAGE_MB=16
run plain LBGPU_STAGE_PLAIN=1
run threads4 LBGPU_COPY_THREADS=4
