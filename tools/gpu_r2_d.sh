#!/bin/bash
# round 2, visit D (2 GPUs): the GPU suite incl. the two-process tests (peer halo, NCCL halo, device init per slab),
# output-path + shim tests; 2-GPU bench: peer vs NCCL halo
mkdir -p gpurun_out/r2d
nvidia-smi topo -m > gpurun_out/r2d/topo.txt 2>&1
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/r2d/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2d/pytest.log | tail -30
for mode in peer nccl; do
  if [ $mode = nccl ]; then export LBGPU_PEER_HALO=0; else unset LBGPU_PEER_HALO; fi
  LBGPU_TRACE=1 LBGPU_VERBOSE=1 timeout 900 python bench.py --gpus 2 --steps 300 --warmup 5 > gpurun_out/r2d/bench_n2_$mode.json 2> gpurun_out/r2d/bench_n2_$mode.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d/bench_n2_$mode.json").read().strip().splitlines()[-1])
    print("$mode n=2", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], d["config"].get("halo"), d["config"]["init_s"], d["e2e"]["kind"], "%.0f" % d["e2e"]["value"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/r2d/bench_n2_$mode.err").read()[-1500:])
PY
done
unset LBGPU_PEER_HALO
LBGPU_TRACE=1 timeout 300 python bench.py --gpus 1 --steps 300 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r2d/bench_n1.json 2> gpurun_out/r2d/bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2d/bench_n1.json").read().strip().splitlines()[-1])
print("n=1", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel ms %.4f" % d["roofline"]["kernel_ms"], "e2e %.0f" % d["e2e"]["value"], d["e2e"]["init_ms"], d["e2e"]["fetch_fields_ms"])
PY
grep "peer halo" gpurun_out/r2d/*.err | head -4
