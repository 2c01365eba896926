#!/bin/bash
# round 2: the bench on 4 GPUs (cfg2 weak scaling + cfg5 strong scaling with the phase trace)
mkdir -p gpurun_out/r2c4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2c4/bench_n4.json 2> gpurun_out/r2c4/bench_n4.err
echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2c4/bench_n4.json").read().strip().splitlines()[-1])
print("n=4", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], "e2e %.0f" % d["e2e"]["value"], "init_s", d["config"]["init_s"])
for k, v in d.get("extra", {}).items():
    print("   ", k, {kk: v[kk] for kk in v if kk not in ("workload", "roofline")})
PY
