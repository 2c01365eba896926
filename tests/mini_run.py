"""Tiny driver for compute-sanitizer: python tests/mini_run.py CASE STEPS"""
import sys
import common, golden_util as gu
name = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = gu.Golden(name)
from hybird_b200 import LB
lb = LB(dict(g.params)); lb.latticeBolzmannInit(*g.init_arrays()); g.configure(lb)
k = 0
for s, F, M, V, W in gu.replay(g, lb, None):
    k += 1
    if k >= steps: break
lb.fetch(); lb.close(); print("ok", name)
