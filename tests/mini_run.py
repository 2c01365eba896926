"""Tiny driver for compute-sanitizer: python tests/mini_run.py CASE STEPS [--dem] [--run]
--dem: the coupled cycle with the DEM sub-steps on the device (lbGpuRunDem); --run: lbGpuRun (graph replay) after the replay"""
import sys
import common, golden_util as gu
name = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = gu.Golden(name)
from hybird_b200 import LB
lb = LB(dict(g.params)); lb.latticeBolzmannInit(*g.init_arrays()); g.configure(lb)
if "--dem" in sys.argv:
    lb.demInit(g.dem())
    for _ in range(steps):
        lb.runDem(1)
    lb.demState(); lb.demContacts()
else:
    k = 0
    for s, F, M, V, W in gu.replay(g, lb, None):
        k += 1
        if k >= steps: break
    if "--run" in sys.argv:
        lb.run(12); lb.synchronize()
lb.fetch(); lb.close(); print("ok", name)
