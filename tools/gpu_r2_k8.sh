#!/bin/bash
# round 2: the bench on 8 GPUs (cfg2 weak scaling + cfg5 strong scaling in the extra block), as the round-end driver launches it
mkdir -p gpurun_out/r2k8
nvidia-smi topo -m > gpurun_out/r2k8/topo.txt 2>&1
LBGPU_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2k8/bench_n8.json 2> gpurun_out/r2k8/bench_n8.err
echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2k8/bench_n8.json").read().strip().splitlines()[-1])
    print("n=8", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], d["config"].get("halo"), "init_s", d["config"]["init_s"], "e2e", d["e2e"]["kind"], "%.0f" % d["e2e"]["value"], d["e2e"]["init_ms"])
    for k, v in d.get("extra", {}).items():
        print("   ", k, {kk: v[kk] for kk in v if kk not in ("roofline", "workload")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2k8/bench_n8.err").read()[-3000:])
PY
grep -c "peer halo on" gpurun_out/r2k8/bench_n8.err
