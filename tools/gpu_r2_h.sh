#!/bin/bash
# round 2, visit H (2 GPUs): GPU suite, then cfg2 + cfg5 on 2 GPUs with the in-cycle exchanges over peer memory / over NCCL
mkdir -p gpurun_out/r2h
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r2h/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2h/pytest.log | tail -30
for mode in peer nccl_cycle; do
  if [ $mode = nccl_cycle ]; then export LBGPU_PEER_CYCLE=0; else unset LBGPU_PEER_CYCLE; fi
  timeout 900 python bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2h/bench_n2_$mode.json 2> gpurun_out/r2h/bench_n2_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2h/bench_n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode n=2", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], d["config"].get("halo"))
    for k, v in d.get("extra", {}).items():
        print("   ", k, {kk: v[kk] for kk in v if kk not in ("roofline", "workload")})
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/r2h/bench_n2_$mode.err").read()[-1500:])
PY
done
