#!/bin/bash
# round 2, visit Y (final profiles of the round) (1 GPU): compute-sanitizer memcheck on three minis, launch lists + ncu full captures for profiles/, bench lines
mkdir -p gpurun_out/r2y
cd tests
for c in cfg5_mini couette_dyn drum_mini; do
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python mini_run.py $c 6 > ../gpurun_out/r2y/sanitizer_memcheck_$c.log 2>&1
  echo "memcheck $c rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2y/sanitizer_memcheck_$c.log | tr '\n' ' ')"
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python mini_run.py bed_dem 6 --dem > ../gpurun_out/r2y/sanitizer_memcheck_bed_dem.log 2>&1
echo "memcheck bed_dem --dem rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2y/sanitizer_memcheck_bed_dem.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python mini_run.py bed_dem 6 --dem > ../gpurun_out/r2y/sanitizer_racecheck_bed_dem.log 2>&1
echo "racecheck bed_dem --dem rc=$? : $(grep -E 'RACECHECK SUMMARY|^ok' ../gpurun_out/r2y/sanitizer_racecheck_bed_dem.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python mini_run.py cfg4_mini 6 --run > ../gpurun_out/r2y/sanitizer_memcheck_cfg4_mini_graph.log 2>&1
echo "memcheck cfg4_mini --run rc=$? : $(grep -E 'ERROR SUMMARY|^ok' ../gpurun_out/r2y/sanitizer_memcheck_cfg4_mini_graph.log | tr '\n' ' ')"
cd ..
for w in cfg3 cfg4 cfg5; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 500 --csv --log-file gpurun_out/r2y/launches_$w.csv python bench.py --workload $w --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y/b_ncu_$w.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y/b_ncu_cfg2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 2 -f -o gpurun_out/r2y/prof_step_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/r2y/prof_step_cfg4 python bench.py --workload cfg4 --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y/b_ncu4f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 3 -f -o gpurun_out/r2y/prof_step_cfg5 python bench.py --workload cfg5 --steps 12 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y/b_ncu5f.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y/bench20.json 2> gpurun_out/r2y/bench20.err
timeout 900 python bench.py --steps 1000 --warmup 10 --no-extra > gpurun_out/r2y/bench1000.json 2> gpurun_out/r2y/bench1000.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2y/bench_ref.json 2> gpurun_out/r2y/bench_ref.err
for w in cfg1 cfg3 cfg4 cfg5; do timeout 600 python bench.py --workload $w --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/r2y/bench_$w.json 2> gpurun_out/r2y/bench_$w.err; done
python - <<PY
import json
for f in ("bench20", "bench1000", "bench_ref", "bench_cfg1", "bench_cfg3", "bench_cfg4", "bench_cfg5"):
    try:
        d = json.loads(open("gpurun_out/r2y/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "MLUPS %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac", d.get("roofline", {}).get("frac"), "e2e", d["e2e"]["value"], d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
