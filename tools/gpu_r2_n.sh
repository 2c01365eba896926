#!/bin/bash
# round 2, visit N (4 GPUs): bench at 4 GPUs (cfg2 weak + cfg5 strong with the phase trace) after the face-launch fix
mkdir -p gpurun_out/r2n
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2n/bench_n4.json 2> gpurun_out/r2n/bench_n4.err
echo "bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2n/bench_n2.json 2> gpurun_out/r2n/bench_n2.err
echo "bench rc=$?"
python - <<PY
import json
for n in (4, 2):
  try:
    d = json.loads(open("gpurun_out/r2n/bench_n%d.json" % n).read().strip().splitlines()[-1])
    print("n=%d" % n, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], "e2e %.0f" % d["e2e"]["value"])
    for k, v in d.get("extra", {}).items():
        print("   ", k, {kk: v[kk] for kk in v if kk not in ("workload",)})
  except Exception as e:
    print("failed", e); print(open("gpurun_out/r2n/bench_n%d.err" % n).read()[-3000:])
PY
