#!/bin/bash
# round 2, visit B (1 GPU): sphere_fixed divergence, whole GPU suite, pipelined-kernel A/B, transfer trace
mkdir -p gpurun_out/r2b
cd tests && timeout 300 python golden_divergence.py sphere_fixed 3 > ../gpurun_out/r2b/sphere_fixed.log 2>&1; cd ..
tail -12 gpurun_out/r2b/sphere_fixed.log
( time timeout 1800 python -m pytest tests -q -m gpu ) > gpurun_out/r2b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b/pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2b/pytest.log | tail -30
for v in "0:" "2:_p2" "3:" "4:_p4" "5:_p5"; do
  pipe=${v%%:*}; suf=${v##*:}
  LBGPU_PIPE=$pipe LBGPU_LIB=$PWD/hybird_b200/liblbgpu$suf.so timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2b/ab_pipe$pipe.json 2> gpurun_out/r2b/ab_pipe$pipe.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2b/ab_pipe$pipe.json").read().strip().splitlines()[-1])
print("pipe=$pipe", "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel frac %.4f" % d["roofline"]["frac"], d["clocks"])
PY
done
LBGPU_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r2b/bench20.json 2> gpurun_out/r2b/bench20.err
grep "lbgpu trace" gpurun_out/r2b/bench20.err | tail -26
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b/bench20.json").read().strip().splitlines()[-1])
print("bench20", "MLUPS %.0f" % d["value"], "frac %.4f" % d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["init_ms"], d["e2e"]["fetch_fields_ms"], d["clocks"])
PY
