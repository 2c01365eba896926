#!/bin/bash
# round 2, visit I2 (1 GPU): k_fs_mass with its loads hoisted: parity (free-surface cases), then cfg4 / cfg5 / cfg1
mkdir -p gpurun_out/r2i2
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_long_full.py -x -q -m gpu > gpurun_out/r2i2/pytest.log 2>&1
echo "rc=$?"; tail -n 4 gpurun_out/r2i2/pytest.log
for w in cfg4 cfg5 cfg1; do
  timeout 600 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r2i2/bench_$w.json 2> gpurun_out/r2i2/bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_fs_mass -s 10 -c 6 --csv --log-file gpurun_out/r2i2/ncu_fsmass_cfg5.csv python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2i2/b_ncu.log 2>&1
grep k_fs_mass gpurun_out/r2i2/ncu_fsmass_cfg5.csv | tail -n 3 | cut -d, -f5,12-
python - <<PY
import json
for w in ("cfg4", "cfg5", "cfg1"):
    d = json.loads(open("gpurun_out/r2i2/bench_%s.json" % w).read().strip().splitlines()[-1])
    print(w, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %.4f" % d["roofline"]["kernel_ms"], "whole %.3f" % d["roofline"]["whole_step_frac"])
PY
