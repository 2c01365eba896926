#!/bin/bash
# round 2, visit M (2 GPUs): the two-process tests (slab cases, device-side DEM on two ranks), bench at 2 GPUs with the phase trace
mkdir -p gpurun_out/r2b2
( time timeout 1500 python -m pytest tests -q -m gpu -k "two_processes or two_gpus or slabs_of_one_device" ) > gpurun_out/r2b2/pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b2/pytest_2gpu.log
tail -n 12 gpurun_out/r2b2/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2b2/bench_n2.json 2> gpurun_out/r2b2/bench_n2.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2b2/bench_n2.json").read().strip().splitlines()[-1])
    print("n=2", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], "e2e %.0f" % d["e2e"]["value"])
    for k, v in d.get("extra", {}).items():
        print("   ", k, {kk: v[kk] for kk in v if kk not in ("roofline", "workload")})
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2b2/bench_n2.err").read()[-3000:])
PY
