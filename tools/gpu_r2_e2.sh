#!/bin/bash
# round 2: peer warm-up in lbGpuCommInit: bench at 2 GPUs (init_s of the first engine), one two-process test
mkdir -p gpurun_out/r2e2
timeout 600 python -m pytest tests -q -m gpu -k "two_processes and cfg5_mini" > gpurun_out/r2e2/pytest.log 2>&1; tail -n 3 gpurun_out/r2e2/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra > gpurun_out/r2e2/bench_n2.json 2> gpurun_out/r2e2/bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2e2/bench_n2.json").read().strip().splitlines()[-1])
print("n=2", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "init_s", d["config"]["init_s"], "e2e %.0f" % d["e2e"]["value"])
PY
